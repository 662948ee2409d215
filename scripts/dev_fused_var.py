"""dev: per-step times of the fused (bg_movegen_eval) and unfused headline step, to see run-to-run structure."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlp_ppo_2ply_multi_b200 as bg
from bench import H, expand_rolls, make_positions, packed_random_weights

dev = torch.device("cuda:0")
w = bg.prepare_weights(packed_random_weights(0).to(dev), H)
boards, players = make_positions(bg, 1048576, dev, seed=2026)
ib, ip, ir = expand_rolls(bg, boards, players)
B = ib.shape[0]
cap = B * 26 + (1 << 20)
pool = torch.empty((cap, 52), dtype=torch.int8, device=dev)
fl = torch.empty(cap, dtype=torch.uint8, device=dev)
vals = torch.empty(cap, dtype=torch.float32, device=dev)
ws = torch.empty(bg._lib.lib().bg_movegen_workspace_bytes(B), dtype=torch.uint8, device=dev)


def fused():
    bg.movegen_evaluate(ib, ip, ir, w, pool, fl, vals, workspace=ws)


def unfused():
    r = bg.movegen(ib, ip, ir, out_boards=pool, out_flags=fl, workspace=ws, want_owner=False, check_status=False)
    bg.evaluate(pool, r.flags, w, n_dev=r.total_dev, out=vals)


for name, fn in (("fused", fused), ("unfused", unfused), ("fused", fused), ("unfused", unfused)):
    ts = []
    for _ in range(12):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(name, " ".join(f"{t:6.1f}" for t in ts))
