"""Profiling driver: position-major move generation (+ evaluation) at a given number of positions.
    python scripts/prof_movegen21.py [positions] [reps] [mode: gen|fused|compact]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mlp_ppo_2ply_multi_b200 as bg
from bench import H, make_positions, packed_random_weights


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    mode = sys.argv[3] if len(sys.argv) > 3 else "gen"
    dev = torch.device("cuda:0")
    boards, players = make_positions(bg, n, dev, 2026)
    P = boards.shape[0]
    pool_cap = P * 21 * 26 + (1 << 20)
    pool = torch.empty((pool_cap, 52), dtype=torch.int8, device=dev)
    flags = torch.empty(pool_cap, dtype=torch.uint8, device=dev)
    codes = torch.empty(pool_cap, dtype=torch.int64, device=dev)
    values = torch.empty(pool_cap, dtype=torch.float32, device=dev)
    ws = torch.empty(bg._lib.lib().bg_movegen_workspace_bytes(21 * P), dtype=torch.uint8, device=dev)
    w = bg.prepare_weights(packed_random_weights(0).to(dev), H)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(reps):
        if mode == "gen":
            r = bg.movegen_all_rolls(boards, players, item_cap=500, out_boards=pool, workspace=ws, out_flags=flags, check_status=False)
        elif mode == "compact":  # the bench step: compact pool, afterstates rebuilt inside the evaluator
            r, _ = bg.movegen_all_rolls_compact(boards, players, w, item_cap=500, out_codes=codes, out_values=values, workspace=ws)
        else:
            r, _ = bg.movegen_evaluate_all_rolls(boards, players, w, pool, flags, values, workspace=ws, item_cap=500)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("positions", P, "afterstates", r.total)


if __name__ == "__main__":
    main()
