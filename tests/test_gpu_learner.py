"""GPU tests of the TD(0) learner (C ABI bg_learner_*, csrc/learner.cu) against the reference Trainer's golden outputs
(tests/golden/learner.npz) and the oracle restatement, and of the reference-facing Trainer surface."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
W_TOL = 1e-5  # abs on every weight after 200 / 400 sequential optimiser steps


@pytest.fixture(scope="module")
def bg():
    import mlp_ppo_2ply_multi_b200 as m

    assert torch.cuda.is_available()
    return m


def _dev(*arrs):
    return [torch.from_numpy(np.ascontiguousarray(a)).to(DEV) for a in arrs]


def test_learner_matches_reference_trainer_golden(bg, golden, oracle):
    g = golden("learner")
    H = int(g["H"])
    L = bg.TD0Learner(H, DEV, lr=float(g["lr"]), gamma=float(g["gamma"]), grad_clip=float(g["grad_clip"]))
    L.set_parameters(torch.from_numpy(g["packed0"]), reset_optimizer=True)
    ob, of, rw, off = _dev(g["obs_boards"], g["obs_flags"], g["reward"], g["ep_offsets"])
    O = oracle.Learner(g["packed0"], H)
    for k in (1, 2):
        met = L.update(ob, of, rw, off).cpu().numpy()
        p = L.packed().cpu().numpy()
        m, v, step = L.optimizer_state()
        assert step == 200 * k
        assert np.abs(p - g[f"packed_after{k}"]).max() < W_TOL  # vs the unmodified reference Trainer
        assert np.abs(m.cpu().numpy() - g[f"m_after{k}"]).max() < 1e-6
        assert np.abs(v.cpu().numpy() - g[f"v_after{k}"]).max() < 1e-7
        assert np.allclose(met.astype(np.float64).mean(0), g["logged"][k - 1], rtol=5e-5, atol=1e-6)
        omet = O.update(g["obs_boards"], g["obs_flags"], g["reward"], g["ep_offsets"])
        assert np.abs(p - O.state()[0]).max() < W_TOL  # vs the oracle
        assert np.allclose(met, omet, rtol=2e-3, atol=2e-6)  # per-episode metrics


def test_learner_is_deterministic_and_records_mode_equals_explicit_observations(bg, golden):
    g = golden("learner")
    H = int(g["H"])
    ob, of, rw, off = _dev(g["obs_boards"], g["obs_flags"], g["reward"], g["ep_offsets"])
    # arena record layout: after_boards[t] = observation board of t+1 within an episode (last one unused by the learner)
    offs = g["ep_offsets"]
    after = np.zeros_like(g["obs_boards"])
    for e in range(len(offs) - 1):
        after[offs[e]:offs[e + 1] - 1] = g["obs_boards"][offs[e] + 1:offs[e + 1]]
        after[offs[e + 1] - 1] = 7  # garbage: must never be read
    meta = (g["obs_flags"] | 0xF4).astype(np.uint8)  # only bit 0 may matter
    ab, mt = _dev(after, meta)
    outs = []
    for mode in range(3):
        L = bg.TD0Learner(H, DEV)
        L.set_parameters(torch.from_numpy(g["packed0"]), reset_optimizer=True)
        met = L.update(ab, mt, rw, off, records=True) if mode == 2 else L.update(ob, of, rw, off)
        outs.append((L.packed().cpu().numpy(), met.cpu().numpy()))
        L.close()
    for p, met in outs[1:]:
        assert np.array_equal(p, outs[0][0]) and np.array_equal(met, outs[0][1])  # bit-identical


@pytest.mark.parametrize("H", [32, 64, 96, 160, 256])
def test_learner_other_hidden_sizes_vs_oracle(bg, golden, oracle, H):
    g = golden("learner")
    rng = np.random.default_rng(H)
    packed = (rng.standard_normal(200 * H + 1) * 0.2).astype(np.float32)
    n_ep = 24
    off = g["ep_offsets"][:n_ep + 1]
    N = int(off[-1])
    ob, of, rw, offd = _dev(g["obs_boards"][:N], g["obs_flags"][:N], g["reward"][:N], off)
    L = bg.TD0Learner(H, DEV)
    L.set_parameters(torch.from_numpy(packed), reset_optimizer=True)
    met = L.update(ob, of, rw, offd).cpu().numpy()
    O = oracle.Learner(packed, H)
    omet = O.update(g["obs_boards"][:N], g["obs_flags"][:N], g["reward"][:N], off)
    # Adam's first steps move a weight by lr * g / (|g| + eps): a gradient that cancels to ~1e-8 turns an fp32-vs-double summation
    # difference into a visible step difference on that one weight, so with random N(0, 0.2) nets the bound is on almost all weights
    d = np.abs(L.packed().cpu().numpy() - O.state()[0])
    assert d.max() < 1e-4 and (d > W_TOL).sum() <= 8
    assert np.allclose(met, omet, rtol=2e-3, atol=2e-6)


def test_learner_explicit_episode_lengths_equal_csr(bg, golden):
    """ep_len: episodes given as (first row, length) in padded, non-contiguous segments (the sync-free all-gather layout) train exactly
    like the same episodes in CSR form; zero-length slots take no optimiser step."""
    g = golden("learner")
    H = int(g["H"])
    off = g["ep_offsets"]
    N, half = int(off[-1]), int(off[100])
    PAD = 12000  # rows per segment
    boards = np.full((2 * PAD, 52), 9, np.int8)
    flags = np.full(2 * PAD, 255, np.uint8)
    rew = np.full(2 * PAD, 123.0, np.float32)
    boards[:half], flags[:half], rew[:half] = g["obs_boards"][:half], g["obs_flags"][:half], g["reward"][:half]
    boards[PAD:PAD + N - half], flags[PAD:PAD + N - half], rew[PAD:PAD + N - half] = g["obs_boards"][half:], g["obs_flags"][half:], g["reward"][half:]
    begin = np.concatenate([off[:100], [half, half], off[100:200] - half + PAD, [0]]).astype(np.int64)  # two empty slots in the middle
    lens = np.concatenate([np.diff(off)[:100], [0, 0], np.diff(off)[100:]]).astype(np.int32)
    outs = []
    for padded in (False, True):
        L = bg.TD0Learner(H, DEV)
        L.set_parameters(torch.from_numpy(g["packed0"]), reset_optimizer=True)
        if padded:
            met = L.update(*_dev(boards, flags, rew, begin), n_episodes=202, ep_len=torch.from_numpy(lens).to(DEV))
            met = met[torch.from_numpy(lens != 0).to(DEV)]
        else:
            met = L.update(*_dev(g["obs_boards"], g["obs_flags"], g["reward"], off))
        assert L.optimizer_state()[2] == 200
        outs.append((L.packed().cpu().numpy(), met.cpu().numpy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("H", [128, 256])
def test_learner_long_episodes_all_tile_counts(bg, golden, oracle, H):
    """episode lengths around the kernel's row-tile (128) and capacity (320) boundaries: 1, 2 and 3 tiles, the longest accepted episode"""
    g = golden("learner")
    lens = [300, 320, 257, 256, 129, 128, 127, 17, 16, 15, 1]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    N = int(off[-1])
    rng = np.random.default_rng(5)
    packed = (rng.standard_normal(200 * H + 1) * 0.1).astype(np.float32)
    boards, flags, rew = g["obs_boards"][:N], g["obs_flags"][:N], g["reward"][:N]
    L = bg.TD0Learner(H, DEV)
    L.set_parameters(torch.from_numpy(packed), reset_optimizer=True)
    met = L.update(*_dev(boards, flags, rew, off)).cpu().numpy()
    O = oracle.Learner(packed, H)
    omet = O.update(boards, flags, rew, off)
    assert L.optimizer_state()[2] == len(lens) and met[:, 5].tolist() == [float(x) for x in lens]
    d = np.abs(L.packed().cpu().numpy() - O.state()[0])
    assert d.max() < 1e-4 and (d > W_TOL).sum() <= 8
    assert np.allclose(met, omet, rtol=2e-3, atol=2e-6)


def test_learner_cuda_core_fallback_kernel(bg, golden, oracle, monkeypatch):
    """H > 128 normally runs the tensor-core kernel in a 10-16 CTA cluster; the 32-units-per-CTA CUDA-core kernel is the fallback when
    the device refuses that cluster size.  Force it and check it against the oracle."""
    monkeypatch.setenv("BG_LEARNER_PATH", "cuda-core")
    test_learner_other_hidden_sizes_vs_oracle(bg, golden, oracle, 256)
    test_learner_other_hidden_sizes_vs_oracle(bg, golden, oracle, 160)


def test_learner_edge_cases(bg, golden, oracle):
    g = golden("learner")
    H = int(g["H"])
    # one-experience episode, empty episode, 3-experience episode, and an episode longer than the kernel accepts
    long_T = 321
    N = 4 + long_T
    boards = np.concatenate([g["obs_boards"][:4], np.repeat(g["obs_boards"][5:6], long_T, 0)])
    flags = np.concatenate([g["obs_flags"][:4], np.zeros(long_T, np.uint8)])
    rew = np.zeros(N, np.float32)
    rew[0], rew[3] = 1.0, 2.0
    off = np.array([0, 1, 1, 4, N], np.int64)
    ob, of, rw, offd = _dev(boards, flags, rew, off)
    L = bg.TD0Learner(H, DEV)
    L.set_parameters(torch.from_numpy(g["packed0"]), reset_optimizer=True)
    with pytest.raises(RuntimeError):
        L.update(ob, of, rw, offd)
    m, v, step = L.optimizer_state()
    assert step == 2  # the empty and the over-long episode took no optimiser step
    O = oracle.Learner(g["packed0"], H)
    O.update(boards[:4], flags[:4], rew[:4], off[:4])
    assert np.abs(L.packed().cpu().numpy() - O.state()[0]).max() < 1e-6
    met = L.update(ob, of, rw, offd[:1])  # zero episodes: a no-op
    assert met.shape == (0, 6) and L.optimizer_state()[2] == 2


def test_trainer_surface_with_reference_format_episodes(bg, golden):
    g = golden("learner")
    H = int(g["H"])
    pm = bg.ParameterManager(hidden_size=H)
    pm.set_parameters(bg.unpack_weights(torch.from_numpy(g["packed0"]), H))
    v0 = pm.get_version()
    tr = bg.Trainer(pm, device=DEV)
    feats = bg.encode(*_dev(g["obs_boards"], g["obs_flags"]))
    off = g["ep_offsets"]
    wt = {0: None, 1: "regular", 2: "gammon", 3: "backgammon"}
    episodes = []
    for e in range(200):
        ep = bg.Episode()
        ep.experiences = [bg.Experience(feats[t], 0.0, torch.tensor(float(g["reward"][t]), device=DEV), t == off[e + 1] - 1, feats[t], 0.0)
                          for t in range(off[e], off[e + 1])]
        ep.win_type = wt[int(g["win_types"][e])]
        episodes.append(ep)
    with pytest.raises(ValueError):
        tr.update(episodes[:199])  # reference trainer.py:49-52
    out = tr.update(episodes)
    assert pm.get_version() == v0 + 1 and tr.total_episodes == 200
    packed = bg.pack_weights(pm.get_parameters()).numpy()
    assert np.abs(packed - g["packed_after1"]).max() < W_TOL
    tags = ["Loss/Training Loss", "TD Error/Mean TD Error", "Gradients/Gradient Norm", "Values/Average Predicted Value",
            "Rewards/Average Reward per Episode", "Episode/Average Episode Length"]
    assert np.allclose([out[t] for t in tags], g["logged"][0], rtol=5e-5, atol=1e-6)
    assert [out["Wins"][k] for k in ("regular", "gammon", "backgammon")] == g["logged_wins"][0].tolist()


def test_arena_to_trainer_loop_zero_copy(bg, golden):
    """config 5 in miniature: arena self-play -> drain 200 episodes -> Trainer.update(EpisodeBatch) -> weights published back."""
    g = golden("learner")
    H = int(g["H"])
    pm = bg.ParameterManager(hidden_size=H)
    pm.set_parameters(bg.unpack_weights(torch.from_numpy(g["packed0"]), H))
    ar = bg.Arena(2048, hidden_size=H, device=DEV, seed=3)
    pm.subscribe(ar)
    tr = bg.Trainer(pm, device=DEV)
    ar.reset()
    while ar.stats()["games"] < 200:
        ar.step(40)
    batch = ar.drain(max_episodes=200)
    assert batch.n_episodes == 200
    # the same batch through explicit observations must give the identical update
    ref = bg.TD0Learner(H, DEV)
    ref.set_parameters(torch.from_numpy(g["packed0"]), reset_optimizer=True)
    ob, of = batch.observation_boards()
    N = batch.n_experiences
    ref.update(ob[:N].contiguous(), of[:N].contiguous(), batch.reward[:N].contiguous(), batch.ep_offsets[:201].contiguous())
    ver = ar.version
    out = tr.update(batch)
    assert torch.equal(tr.learner.packed(), ref.packed())
    assert ar.version == ver + 1 and torch.equal(ar._packed, tr.learner.packed())  # arena now plays with the new weights
    assert out["Episode/Average Episode Length"] == pytest.approx(N / 200)
    assert sum(out["Wins"].values()) <= 200
    ar.step(5)
    assert ar.stats()["errors"] == 0
    ar.close()
