"""Timing driver: 1-ply self-play, G concurrent games (BASELINE configs[2]); ms per ply and games/s."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mlp_ppo_2ply_multi_b200 as bg
from bench import H, packed_random_weights

G = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
plies = int(sys.argv[2]) if len(sys.argv) > 2 else 200
dev = torch.device("cuda:0")
packed = packed_random_weights(0).to(dev)
ar = bg.Arena(G, hidden_size=H, device=dev, seed=0, ring_experiences=G * 48, ring_episodes=G)
ar.set_weights(packed, version=1)
ar.reset()
ar.step(120)
ar.drain(max_episodes=G, max_experiences=G * 48)
torch.cuda.synchronize()
s0 = ar.stats()
a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a0.record()
done = 0
while done < plies:
    ar.step(20)
    ar.drain(max_episodes=G, max_experiences=G * 48)
    done += 20
a1.record()
torch.cuda.synchronize()
s1 = ar.stats()
ms = a0.elapsed_time(a1)
print(f"{G} games: {ms / done:.4f} ms per ply, {(s1['games'] - s0['games']) / ms * 1e3:,.0f} games/s, {(s1['afterstates'] - s0['afterstates']) / ms / 1e6:.2f} G afterstates/s, errors {s1['errors']}")
