"""dev: 1-ply self-play throughput with K arenas of G/K games on K streams (tails of one overlap the bulk of the others)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlp_ppo_2ply_multi_b200 as bg
from bench import packed_random_weights

dev = torch.device("cuda:0")
H = 128
packed = packed_random_weights(0).to(dev)
G = 65536
for K in (1, 2, 4):
    n = G // K
    ars = [bg.Arena(n, hidden_size=H, device=dev, seed=0, game_id_base=k * n, ring_experiences=n * 48, ring_episodes=n) for k in range(K)]
    sts = [torch.cuda.Stream(device=dev) for _ in range(K)]
    for ar, s in zip(ars, sts):
        with torch.cuda.stream(s):
            ar.set_weights(packed, version=1)
            ar.reset()
            ar.step(120)
            ar.drain(max_episodes=n, max_experiences=n * 48)
    torch.cuda.synchronize()
    g0 = sum(ar.stats()["games"] for ar in ars)
    t0 = time.perf_counter()
    plies = 0
    while plies < 200:
        for ar, s in zip(ars, sts):
            with torch.cuda.stream(s):
                ar.step(20)
        for ar, s in zip(ars, sts):
            with torch.cuda.stream(s):
                ar.drain(max_episodes=n, max_experiences=n * 48)
        plies += 20
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    g1 = sum(ar.stats()["games"] for ar in ars)
    print(f"K={K}: {(g1 - g0) / dt:,.0f} games/s, {dt / plies * 1e3:.3f} ms per ply of all {G} games")
    for ar in ars:
        ar.close()
