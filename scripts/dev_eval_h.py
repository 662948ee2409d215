"""dev: evaluator throughput by hidden size (generic FFMA kernel vs the H = 128 kernels)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlp_ppo_2ply_multi_b200 as bg
from bench import make_positions

dev = torch.device("cuda:0")
boards, players = make_positions(bg, 1 << 20, dev, seed=3)
boards = boards.repeat(8, 1).contiguous()
flags = players.repeat(8).contiguous()
N = boards.shape[0]
for H in (32, 64, 128, 256):
    w = bg.prepare_weights((torch.randn(200 * H + 1) * 0.1).to(dev), H)
    out = torch.empty(N, device=dev)
    for _ in range(2):
        bg.evaluate(boards, flags, w, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        bg.evaluate(boards, flags, w, out=out)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"H={H}: {N / ms / 1e6:.2f} G boards/s ({ms:.2f} ms for {N} boards)")
