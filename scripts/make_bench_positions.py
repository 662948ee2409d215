"""Writes tests/golden/bench_positions.npz: the first 65,536 positions of bench.py's config-2 position set (the arena playing uniformly
random legal moves from the start position, snapshotted at spread-out plies; bench.make_positions, seed 2026).  Needs a GPU; run once:
    python scripts/make_bench_positions.py gpurun_out/bench_positions.npz     (then copy the file into tests/golden/)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mlp_ppo_2ply_multi_b200 as bg
from bench import make_positions

out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "bench_positions.npz")
b, p = make_positions(bg, 65536, torch.device("cuda:0"), 2026)
np.savez_compressed(out, boards=b.cpu().numpy(), players=p.cpu().numpy())
print("wrote", out, b.shape, "P1 to move:", int((p == 0).sum().item()))
