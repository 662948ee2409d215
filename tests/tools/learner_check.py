"""dev: max |dW| of bg_learner_update vs the oracle for several hidden sizes."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mlp_ppo_2ply_multi_b200 as bg
from oracle import pyoracle as po

g = np.load(os.path.join(os.path.dirname(__file__), "..", "golden", "learner.npz"))
dev = "cuda:0"
n_ep = 24
off = g["ep_offsets"][:n_ep + 1]
N = int(off[-1])
ob, of, rw, offd = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (g["obs_boards"][:N], g["obs_flags"][:N], g["reward"][:N], off))
for H in (32, 64, 96, 128, 160, 192, 224, 256):
    rng = np.random.default_rng(H)
    packed = (rng.standard_normal(200 * H + 1) * 0.2).astype(np.float32)
    L = bg.TD0Learner(H, dev)
    L.set_parameters(torch.from_numpy(packed), reset_optimizer=True)
    O = po.Learner(packed, H)
    for k in range(1, n_ep + 1):
        sl = slice(int(off[k - 1]), int(off[k]))
        o2 = torch.tensor([0, off[k] - off[k - 1]], dtype=torch.int64, device=dev)
        met = L.update(ob[sl].contiguous(), of[sl].contiguous(), rw[sl].contiguous(), o2).cpu().numpy()
        omet = O.update(g["obs_boards"][sl], g["obs_flags"][sl], g["reward"][sl], np.array([0, off[k] - off[k - 1]]))
        d = np.abs(L.packed().cpu().numpy() - O.state()[0])
        if k in (1, 2, 3, n_ep) or d.max() > 1e-5:
            i = int(d.argmax())
            print(f"H={H} ep {k}: max|dW| {d.max():.3e} at {i} (row {i // H}, unit {i % H})  n>1e-5: {(d > 1e-5).sum()}  met {met[0][:3]} vs {omet[0][:3]}")
            if d.max() > 1e-5:
                break
