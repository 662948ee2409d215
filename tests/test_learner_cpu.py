"""CPU tests of the learner's checker: the oracle restatement of Trainer.update (oracle/bg_oracle.c bgo_learner_*) against
tests/golden/learner.npz, which holds the outputs of the UNMODIFIED reference Trainer (tests/golden/make_golden_learner.py):
two consecutive 200-episode updates = 400 sequential TD(0)/Adam steps.  Plus the host-side feature -> board inverse."""
import numpy as np
import torch

W_TOL = 1e-5  # abs, on every weight after 200 / 400 sequential optimiser steps (fp32 rounding, summation order unspecified in torch)


def test_oracle_learner_matches_reference_trainer(oracle, golden):
    g = golden("learner")
    H = int(g["H"])
    L = oracle.Learner(g["packed0"], H, lr=float(g["lr"]), gamma=float(g["gamma"]), grad_clip=float(g["grad_clip"]))
    for k in (1, 2):
        met = L.update(g["obs_boards"], g["obs_flags"], g["reward"], g["ep_offsets"])
        p, m, v, step = L.state()
        assert step == 200 * k
        assert np.abs(p - g[f"packed_after{k}"]).max() < W_TOL
        assert np.abs(m - g[f"m_after{k}"]).max() < 1e-7
        assert np.abs(v - g[f"v_after{k}"]).max() < 1e-8
        # the six scalars the reference logs (trainer.py:157-163, 195-210) are batch means of the per-episode metrics
        assert np.allclose(met.astype(np.float64).mean(0), g["logged"][k - 1], rtol=2e-5, atol=1e-7)
    wins = np.bincount(g["win_types"], minlength=4)[1:4]
    assert np.array_equal(wins, g["logged_wins"][0])


def test_oracle_learner_single_experience_episode_and_empty_episode(oracle, golden):
    g = golden("learner")
    H = int(g["H"])
    L = oracle.Learner(g["packed0"], H)
    # episode 0: one experience (target = reward, trainer.py:113); episode 1: empty (skipped); episode 2: three experiences
    off = np.array([0, 1, 1, 4], np.int64)
    met = L.update(g["obs_boards"][:4], g["obs_flags"][:4], np.array([1.0, 0.0, 0.0, 2.0], np.float32), off)
    p, m, v, step = L.state()
    assert step == 2 and met[1].tolist() == [0.0] * 6 and met[0][5] == 1 and met[2][5] == 3
    v0 = oracle.value(g["packed0"], H, g["obs_boards"][:1], g["obs_flags"][:1])[0]
    assert abs(met[0][0] - (v0 - 1.0) ** 2) < 1e-6  # loss of the one-row episode


def test_features_to_boards_is_exact_inverse(oracle):
    import mlp_ppo_2ply_multi_b200 as bg

    boards, players = oracle.random_positions(2000, seed=77)
    flags = (np.arange(2000) % 2).astype(np.uint8)
    feats = oracle.encode(boards, flags)
    b2, f2 = bg.features_to_boards(torch.from_numpy(feats))
    assert np.array_equal(b2.numpy(), boards) and np.array_equal(f2.numpy(), flags)
