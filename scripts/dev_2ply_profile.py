import os, sys, torch
sys.path.insert(0, "/root/repo")
import mlp_ppo_2ply_multi_b200 as bg
from bench import packed_random_weights
G = 65536
dev = torch.device("cuda:0")
ar = bg.Arena(G, device=dev, seed=0, ring_experiences=G * 64, ring_episodes=G)
ar.set_weights(packed_random_weights(0).to(dev), version=1)
ar.set_lookahead(4, 5, 1.0, 0.9)
ar.reset(); ar.step(60); ar.drain(max_episodes=G, max_experiences=G * 64)
ar.step(2, lookahead=2); torch.cuda.synchronize()
torch.cuda.profiler.start(); ar.step(1, lookahead=2); torch.cuda.synchronize(); torch.cuda.profiler.stop()
