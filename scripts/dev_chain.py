"""dev: cost of the (mostly empty) tier chain on a tiny per-item batch"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mlp_ppo_2ply_multi_b200 as bg
import numpy as np
dev = torch.device("cuda:0")
b = np.zeros((64, 52), np.int8); b[:, 0] = 2; b[:, 11] = 5; b[:, 16] = 3; b[:, 18] = 5; b[:, 24 + 23] = 2; b[:, 24 + 12] = 5; b[:, 24 + 7] = 3; b[:, 24 + 5] = 5
boards = torch.from_numpy(b).to(dev); players = torch.zeros(64, dtype=torch.uint8, device=dev)
rolls = torch.tensor([[3, 1]] * 64, dtype=torch.uint8, device=dev)
pool = torch.empty((4096, 52), dtype=torch.int8, device=dev)
for _ in range(3):
    r = bg.movegen(boards, players, rolls, out_boards=pool, check_status=False, want_owner=False)
torch.cuda.synchronize()
torch.cuda.profiler.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    r = bg.movegen(boards, players, rolls, out_boards=pool, check_status=False, want_owner=False)
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("per call", e0.elapsed_time(e1) / 50 * 1e3, "us; total", r.total)
