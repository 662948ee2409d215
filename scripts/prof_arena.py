"""Profiling driver: a few 1-ply self-play plies of G games inside the profiler range."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mlp_ppo_2ply_multi_b200 as bg
from bench import H, packed_random_weights

G = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
plies = int(sys.argv[2]) if len(sys.argv) > 2 else 3
look = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda:0")
ar = bg.Arena(G, hidden_size=H, device=dev, seed=0, ring_experiences=G * 48, ring_episodes=G)
ar.set_weights(packed_random_weights(0).to(dev), version=1)
if look == 2:
    ar.set_lookahead(4, 5, 1.0, 0.9)
ar.reset()
ar.step(120)
ar.drain(max_episodes=G, max_experiences=G * 48)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ar.step(plies, lookahead=look)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", ar.stats()["errors"])
