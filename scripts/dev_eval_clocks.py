"""dev: SM clock / power while the tcgen05 evaluator runs back to back for ~3 s on the bench workload (is the kernel power limited?)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mlp_ppo_2ply_multi_b200 as bg  # noqa: E402
from bench import H, ClockSampler, make_positions, packed_random_weights  # noqa: E402

dev = torch.device("cuda:0")
boards, players = make_positions(bg, 1048576, dev, 2026)
w = bg.prepare_weights(packed_random_weights(0).to(dev), H)
r = bg.movegen_all_rolls_compact(boards, players, None, item_cap=500)
out = torch.empty(r.codes.numel(), dtype=torch.float32, device=dev)
for _ in range(3):
    bg.evaluate_codes(r, w, out=out)
torch.cuda.synchronize()
sampler = ClockSampler(0)
time.sleep(0.7)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 80
e0.record()
for _ in range(reps):
    bg.evaluate_codes(r, w, out=out)
e1.record()
torch.cuda.synchronize()
print("eval (compact), back to back:", e0.elapsed_time(e1) / reps, "ms per pass")
print(sampler.stop())
