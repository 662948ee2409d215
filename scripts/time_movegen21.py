"""Timing driver: position-major move generation at a given number of positions (CUDA events, per pass)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mlp_ppo_2ply_multi_b200 as bg
from bench import make_positions


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
    dev = torch.device("cuda:0")
    boards, players = make_positions(bg, n, dev, 2026)
    P = boards.shape[0]
    pool_cap = P * 21 * 26 + (1 << 20)
    pool = torch.empty((pool_cap, 52), dtype=torch.int8, device=dev)
    flags = torch.empty(pool_cap, dtype=torch.uint8, device=dev)
    ws = torch.empty(bg._lib.lib().bg_movegen_workspace_bytes(21 * P), dtype=torch.uint8, device=dev)
    for dbg in sys.argv[2:] or ["1"]:
        os.environ["BG_MG21_DEBUG"] = dbg
        fn = lambda: bg.movegen_all_rolls(boards, players, item_cap=500, out_boards=pool, workspace=ws, out_flags=flags, check_status=False)
        for _ in range(2):
            r = fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        R = 5
        for _ in range(R):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / R
        ovf = ws[32:48].view(torch.int32).cpu().tolist()
        print(f"debug={dbg}: {ms:.2f} ms per pass, total {r.total}, overflow lists {ovf}", flush=True)


if __name__ == "__main__":
    main()
