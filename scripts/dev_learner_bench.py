"""dev: time bg_learner_update on the golden 200-episode batch (us per episode) for both observation modes."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlp_ppo_2ply_multi_b200 as bg

g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "learner.npz"))
dev = "cuda:0"
H = int(g["H"])
ob, of, rw, off = (torch.from_numpy(np.ascontiguousarray(g[k])).to(dev) for k in ("obs_boards", "obs_flags", "reward", "ep_offsets"))
for Hx in (H, 256):
    L = bg.TD0Learner(Hx, dev)
    L.set_parameters(torch.randn(200 * Hx + 1) * 0.1, reset_optimizer=True)
    for _ in range(3):
        L.update(ob, of, rw, off, check_status=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    R = 10
    for _ in range(R):
        L.update(ob, of, rw, off, check_status=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / R
    print(f"H={Hx}: {ms:.3f} ms per 200-episode update ({int(off[-1])} experiences) = {ms * 1e3 / 200:.2f} us/episode, "
          f"{200 / ms * 1e3:.0f} episodes/s, {int(off[-1]) / ms * 1e3:.0f} experiences/s")
