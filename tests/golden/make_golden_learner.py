#!/usr/bin/env python
"""Golden vectors for the learner: outputs of the UNMODIFIED reference Trainer.update
(/root/reference/src/agents/trainer.py:48-166) on a fixed batch of 200 self-play episodes.

Run in the build container only (the reference does not travel to the GPU box):
    python tests/golden/make_golden_learner.py
Output (committed): tests/golden/learner.npz

The episodes are oracle self-play games (the oracle's env/move generator/encoder are pinned bit-exact against the
reference elsewhere); the observations handed to the reference Trainer are built with the REFERENCE encoder
(ImmutableBoard.get_board_features).  Trainer.update is called twice on the same batch so that the second call starts
from non-trivial Adam moments (steps 201-400).  Nothing of the trainer is modified: the NVML handle it asks for and the
S3 logger are stubbed because this container has neither a GPU nor boto3.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from oracle import reference_shim as shim  # noqa: E402

shim.install()
torch.set_num_threads(1)

import pynvml  # noqa: E402


class _Util:
    gpu = 0


class _Mem:
    used = 0


pynvml.nvmlDeviceGetHandleByIndex = lambda i: None
pynvml.nvmlDeviceGetUtilizationRates = lambda h: _Util()
pynvml.nvmlDeviceGetMemoryInfo = lambda h: _Mem()

from agents import BackgammonPolicyNetwork  # noqa: E402
from agents.trainer import Trainer  # noqa: E402
from environments.episode import Episode, Experience  # noqa: E402
from backgammon.types import Player  # noqa: E402


class _PM:
    """the two ParameterManager methods Trainer touches (trainer.py:21,166)"""

    def __init__(self, sd):
        self.sd = sd
        self.n_set = 0

    def get_parameters(self):
        return {k: v.clone() for k, v in self.sd.items()}

    def set_parameters(self, sd):
        self.sd = {k: v.detach().clone() for k, v in sd.items()}
        self.n_set += 1


class _Scalars:
    def __init__(self):
        self.scalars = {}

    def add_scalar(self, tag, value, step):
        self.scalars.setdefault(tag, []).append(float(value))

    def add_scalars(self, tag, d, step):
        for k, v in d.items():
            self.scalars.setdefault(f"{tag}/{k}", []).append(float(v))

    def add_histogram(self, *a, **k):
        pass

    class writer:
        @staticmethod
        def flush():
            pass


def main():
    torch.manual_seed(0)
    net = BackgammonPolicyNetwork()  # H = 128, Xavier (policy_network.py:36-51)
    H = net.fc1.out_features
    packed0 = po.pack_weights(net.state_dict())
    obs_b, obs_f, rew, off, wins = po.selfplay_episodes(packed0, H, 200, temperature=1.5, seed=5)
    assert len(off) == 201
    feats = []
    for b, f in zip(obs_b, obs_f):
        feats.append(shim.array_to_board_env(b).get_board_features(Player(int(f))))
    feats = torch.stack(feats)
    assert np.array_equal(feats.numpy(), po.encode(obs_b, obs_f)), "oracle encoder disagrees with the reference"

    episodes = []
    wt = {0: None, 1: "regular", 2: "gammon", 3: "backgammon"}
    for e in range(200):
        ep = Episode()
        for t in range(off[e], off[e + 1]):
            ep.experiences.append(Experience(observation=feats[t], state_value=0.0, reward=torch.tensor(float(rew[t])), done=bool(t == off[e + 1] - 1),
                                             next_observation=feats[t], next_state_value=0.0))
        ep.win_type = wt[int(wins[e])]
        episodes.append(ep)

    pm = _PM(net.state_dict())
    tr = Trainer(pm, device=torch.device("cpu"))
    tr.logger = _Scalars()
    out = {}
    for k in (1, 2):
        tr.update(episodes)
        out[f"packed_after{k}"] = po.pack_weights(pm.sd)
        st = tr.optimizer.state_dict()["state"]
        # Adam moments in packed order: fc1.weight (transposed), fc1.bias, value_head.weight, value_head.bias
        for name, key in (("m", "exp_avg"), ("v", "exp_avg_sq")):
            out[f"{name}_after{k}"] = np.concatenate([st[0][key].numpy().T.reshape(-1), st[1][key].numpy().reshape(-1), st[2][key].numpy().reshape(-1),
                                                      st[3][key].numpy().reshape(-1)]).astype(np.float32)
    sc = tr.logger.scalars
    tags = ["Loss/Training Loss", "TD Error/Mean TD Error", "Gradients/Gradient Norm", "Values/Average Predicted Value",
            "Rewards/Average Reward per Episode", "Episode/Average Episode Length"]
    out["logged"] = np.array([sc[t] for t in tags], np.float64).T  # [2 updates, 6]
    out["logged_wins"] = np.array([[sc[f"Wins/{w}"][k] for w in ("regular", "gammon", "backgammon")] for k in range(2)], np.int64)
    np.savez_compressed(os.path.join(HERE, "learner.npz"), H=np.int32(H), packed0=packed0, obs_boards=obs_b, obs_flags=obs_f, reward=rew,
                        ep_offsets=off, win_types=wins, lr=np.float32(1e-3), gamma=np.float32(0.99), grad_clip=np.float32(1.0), **out)
    print("episodes", 200, "experiences", int(off[-1]), "logged", out["logged"])

    # oracle restatement vs the reference
    L = po.Learner(packed0, H)
    for k in (1, 2):
        met = L.update(obs_b, obs_f, rew, off)
        p, m, v, step = L.state()
        print(f"update {k}: max|dW| oracle vs reference = {np.abs(p - out[f'packed_after{k}']).max():.3e}  "
              f"max|dm| = {np.abs(m - out[f'm_after{k}']).max():.3e}  max|dv| = {np.abs(v - out[f'v_after{k}']).max():.3e}  "
              f"metrics(mean) = {met.mean(0)}  logged = {out['logged'][k - 1]}")


if __name__ == "__main__":
    main()
