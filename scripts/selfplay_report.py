#!/usr/bin/env python
"""1-ply self-play, 65,536 concurrent games on one B200, run until >= 1,048,576 games finish, for
{shipped 2.1 M-episode checkpoint, seed-0 Xavier init} x {T = 1.5, T = 0.5, greedy}; dice Philox seed 0.  Prints one line per setting:
games/s, plies/s, afterstates/s, mean plies per game, pass rate, win-type mix.
    python scripts/selfplay_report.py > profiles/r01_selfplay_config3_matrix.txt"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mlp_ppo_2ply_multi_b200 as bg  # noqa: E402

dev = torch.device("cuda:0")
g = np.load(os.path.join(ROOT, "tests", "golden", "values.npz"))
H = int(g["H"])
G, TARGET = 65536, 1 << 20
print(f"# {G} concurrent games, until >= {TARGET} finished; afterstates = sum of legal moves over decisions (the current-position row is evaluated separately)")
for wname, key in (("checkpoint backgammon_256_standard_episode_2100000.pth", "packed"), ("Xavier init, torch.manual_seed(0)", "packed_init0")):
    packed = torch.from_numpy(g[key]).to(dev)
    for T in (1.5, 0.5, 0.0):
        ar = bg.Arena(G, hidden_size=H, device=dev, seed=0, ring_experiences=G * 48, ring_episodes=G)
        ar.set_weights(packed, version=1, temperature=T)
        ar.reset()
        for _ in range(6):  # desynchronise game phases before timing (ring drained as it fills)
            ar.step(25)
            while ar.drain(max_episodes=G, max_experiences=G * 48).n_episodes:
                pass
        torch.cuda.synchronize()
        s0 = ar.stats()
        t0 = time.perf_counter()
        plies = 0
        while True:
            ar.step(20)
            ar.drain(max_episodes=G, max_experiences=G * 48)
            plies += 20
            s1 = ar.stats()
            if s1["games"] - s0["games"] >= TARGET:
                break
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        d = {k: s1[k] - s0[k] for k in s1}
        wt = np.array([d["win_regular"], d["win_gammon"], d["win_backgammon"], d["truncated"]], float) / max(d["games"], 1)
        print(f"{wname:55s} T={'greedy' if T == 0 else T}: {d['games'] / dt:11,.0f} games/s  {G * plies / dt:13,.0f} plies/s  {d['afterstates'] / dt:15,.0f} afterstates/s  "
              f"{d['steps'] / d['games']:6.1f} plies/game  passes {100 * d['passes'] / d['steps']:4.1f} %  regular/gammon/backgammon/truncated "
              f"{wt[0]:.3f}/{wt[1]:.3f}/{wt[2]:.3f}/{wt[3]:.4f}  P1 wins {d['p1_wins'] / max(d['games'], 1):.3f}  errors {d['errors']}  wait {d['wait_steps']}", flush=True)
        ar.close()
