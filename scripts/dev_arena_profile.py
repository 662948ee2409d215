"""Developer tool: a few arena plies under `ncu --profile-from-start off` (launch list), and plain per-ply timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mlp_ppo_2ply_multi_b200 as bg
from bench import packed_random_weights

G = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda:0")
ar = bg.Arena(G, device=dev, seed=0, ring_experiences=G * 64, ring_episodes=G)
ar.set_weights(packed_random_weights(0).to(dev), version=1)
ar.reset()
ar.step(120)
ar.drain(max_episodes=G, max_experiences=G * 64)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ar.step(40); e1.record(); torch.cuda.synchronize()
print("ms per ply (no drain):", e0.elapsed_time(e1) / 40)
t = time.perf_counter(); b = ar.drain(max_episodes=G, max_experiences=G * 64); torch.cuda.synchronize(); print("drain ms", (time.perf_counter() - t) * 1e3, b.n_episodes)
if os.environ.get("BG_PROFILE") == "1":
    torch.cuda.profiler.start(); ar.step(3); torch.cuda.synchronize(); torch.cuda.profiler.stop()
print(ar.stats())
