"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: game sharding, the weight-blob broadcast that replaces
ParameterManager polling, and the statistics reduction.  The data path itself has no collective (games are independent)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mlp_ppo_2ply_multi_b200 as bg
    from mlp_ppo_2ply_multi_b200 import distributed as bgd

    torch.manual_seed(100 + rank)  # ranks start from DIFFERENT weights
    pm = bg.ParameterManager(hidden_size=128)

    class FakeArena:  # records what a subscribed arena would receive
        device = torch.device("cpu")

        def __init__(self):
            self.got = []

        def set_weights(self, packed, version, temperature):
            self.got.append((packed.clone(), version, temperature))

    fa = FakeArena()
    pm.subscribe(fa)  # LOCAL (no hidden collective): this rank's own initial weights
    assert len(fa.got) == 1 and fa.got[0][1] == 1
    if rank == 0:
        net = bg.BackgammonPolicyNetwork()
        pm.set_parameters(net.state_dict())  # trainer rank publishes -> one broadcast
        pm.set_parameters(net.state_dict())
    else:
        pm.sync_from_source()
        pm.sync_from_source()
    packed = bg.pack_weights(pm.get_parameters())
    assert len(fa.got) == 3 and fa.got[-1][1] == 3 and torch.equal(fa.got[-1][0], packed) and pm.check_version_sync()
    # the training loop of examples/train_selfplay.py with evaluation enabled: iteration u launches update u-1 on the batch staged one
    # iteration earlier, so rank 0 publishes from iteration 2 on, the other ranks receive at the same point (before the episode gather), and
    # the evaluation branch issues NO collective (it reads the already published weights)
    for u in range(6):
        if rank == 0:
            if u > 1:
                pm.set_packed(packed + u)
        elif u > 1:
            pm.sync_from_source()
        torch.distributed.all_reduce(torch.zeros(1))  # stands for the episode gather
        if rank == 0 and u and u % 2 == 0:
            _ = bg.pack_weights(pm.get_parameters())  # evaluation: no publish, no finish
    if rank == 0:
        pm.set_packed(packed + 5)
    else:
        pm.sync_from_source()
    assert pm.get_version() == 8 and pm.check_version_sync() and torch.equal(bg.pack_weights(pm.get_parameters()), packed + 5)
    assert fa.got[-1][1] == 8 and fa.got[-1][2] == pm.get_temperature()
    pm.set_packed(packed) if rank == 0 else pm.sync_from_source()
    # the reference's constructor arguments (lock, version, parameters) are accepted
    assert bg.ParameterManager(object(), object(), {}).get_version() == 1
    n_local, base = bgd.shard_games(65536 + 1, rank, world)
    stats = bgd.all_reduce_stats({"games": 10 + rank, "steps": 1000 * (rank + 1), "afterstates": 7})
    w, ver, temp = bgd.broadcast_weights(packed * (rank + 1), 5 + rank, 1.25 - rank, src=0)
    # collective 3: every rank's drained episodes gathered into one batch (here: synthetic CPU batches of different sizes)
    from mlp_ppo_2ply_multi_b200.episode import EpisodeBatch

    E, lens = 2 + rank, [3 + rank, 1, 4][: 2 + rank]
    N = sum(lens)
    g = torch.Generator().manual_seed(7 + rank)
    eb = EpisodeBatch(E, N, torch.randint(0, 6, (N + 5, 52), generator=g, dtype=torch.int8), torch.randint(0, 32, (N + 5,), generator=g, dtype=torch.uint8),
                      torch.rand(N + 5, generator=g), torch.rand(N + 5, generator=g), torch.rand(N + 5, generator=g),
                      torch.randint(1, 50, (N + 5,), generator=g, dtype=torch.int16), torch.randint(0, 50, (N + 5,), generator=g, dtype=torch.int16),
                      torch.randint(1, 7, (N + 5, 2), generator=g, dtype=torch.uint8), torch.tensor([0] + list(np.cumsum(lens)) + [0, 0], dtype=torch.int64),
                      torch.full((E + 2, 12), 100 + rank, dtype=torch.int32))
    merged = bgd.all_gather_episodes(eb, max_episodes=4, max_experiences=12)
    padded = bgd.all_gather_episodes(eb, max_episodes=4, max_experiences=12, compact=False)  # no host sync; explicit episode lengths
    slim = bgd.all_gather_episodes(eb, max_episodes=4, max_experiences=12, compact=False, fields=bgd.LEARNER_FIELDS)  # what the trainer reads
    assert slim.roll is None and slim.action is None and torch.equal(slim.after_boards, padded.after_boards) and torch.equal(slim.ep_len, padded.ep_len)
    ob, of = padded.observation_boards()  # zero-length filler episodes must not index out of range
    assert ob.shape == (24, 52)
    np.savez(os.path.join(out_dir, f"pad{rank}.npz"), n=np.array([padded.n_episodes, padded.n_experiences]), begin=padded.ep_offsets.numpy(),
             len=padded.ep_len.numpy(), after=padded.after_boards.numpy(), reward=padded.reward.numpy(), info=padded.ep_info.numpy())
    np.savez(os.path.join(out_dir, f"ep{rank}.npz"), n=np.array([merged.n_episodes, merged.n_experiences]), off=merged.ep_offsets.numpy(),
             after=merged.after_boards.numpy(), reward=merged.reward.numpy(), info=merged.ep_info.numpy(), roll=merged.roll.numpy(),
             action=merged.action.numpy(), my_after=eb.after_boards[:N].numpy(), my_reward=eb.reward[:N].numpy())
    # the padded form hands out TWO buffer sets in turn: a batch stays valid until the second next gather with the same quota (the learner
    # kernel may still be reading batch u while batch u + 1 is gathered)
    slim = bgd.all_gather_episodes(eb, max_episodes=4, max_experiences=12, compact=False, fields=bgd.LEARNER_FIELDS)
    keep = slim.after_boards.clone()
    eb2 = EpisodeBatch(E, N, torch.zeros_like(eb.after_boards), eb.meta, eb.reward, eb.state_value, eb.next_state_value, eb.n_moves, eb.action, eb.roll,
                       eb.ep_offsets, eb.ep_info)
    nxt = bgd.all_gather_episodes(eb2, max_episodes=4, max_experiences=12, compact=False, fields=bgd.LEARNER_FIELDS)
    assert nxt.after_boards.data_ptr() != slim.after_boards.data_ptr() and torch.equal(slim.after_boards, keep) and int(nxt.after_boards.abs().sum()) == 0
    third = bgd.all_gather_episodes(eb, max_episodes=4, max_experiences=12, compact=False, fields=bgd.LEARNER_FIELDS)
    assert third.after_boards.data_ptr() == slim.after_boards.data_ptr() and torch.equal(third.after_boards, keep)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), packed=packed.numpy(), version=pm.get_version(), temperature=pm.get_temperature(),
             n_local=n_local, base=base, games=stats["games"], steps=stats["steps"], after=stats["afterstates"], w=w.numpy(), ver=ver, temp=temp)
    dist.destroy_process_group()


def test_two_rank_weight_broadcast_sharding_and_stats(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (np.load(tmp_path / f"r{r}.npz") for r in range(world))
    assert np.array_equal(r0["packed"], r1["packed"])  # both ranks hold the trainer rank's weights
    assert int(r0["version"]) == int(r1["version"]) == 9
    assert float(r0["temperature"]) == float(r1["temperature"]) == pytest.approx(1.5 - 1.0 * 8 / 4000)
    assert int(r0["n_local"]) + int(r1["n_local"]) == 65537 and int(r0["base"]) == 0 and int(r1["base"]) == int(r0["n_local"])
    assert int(r0["games"]) == int(r1["games"]) == 21 and int(r0["steps"]) == 3000 and int(r1["after"]) == 14
    assert np.array_equal(r0["w"], r1["w"]) and int(r1["ver"]) == 5 and float(r1["temp"]) == 1.25
    e0, e1 = (np.load(tmp_path / f"ep{r}.npz") for r in range(world))
    for k in ("n", "off", "after", "reward", "info", "roll", "action"):
        assert np.array_equal(e0[k], e1[k])  # every rank holds the same merged batch
    assert e0["n"].tolist() == [5, 4 + 9] and e0["off"].tolist() == [0, 3, 4, 8, 9, 13]  # rank 0: episodes of 3, 1; rank 1: 4, 1, 4
    assert np.array_equal(e0["after"], np.concatenate([e0["my_after"], e1["my_after"]]))
    assert np.array_equal(e0["reward"], np.concatenate([e0["my_reward"], e1["my_reward"]]))
    assert e0["info"][:, 0].tolist() == [100, 100, 101, 101, 101]
    p0 = np.load(tmp_path / "pad0.npz")
    assert p0["n"].tolist() == [8, 24] and p0["len"].tolist() == [3, 1, 0, 0, 4, 1, 4, 0]  # world * max_episodes slots, missing ones empty
    k = 0
    for slot in range(8):  # the non-empty slots, in order, are exactly the compact batch's episodes
        if p0["len"][slot]:
            lo, n = int(p0["begin"][slot]), int(p0["len"][slot])
            assert np.array_equal(p0["after"][lo:lo + n], e0["after"][e0["off"][k]:e0["off"][k + 1]])
            assert np.array_equal(p0["reward"][lo:lo + n], e0["reward"][e0["off"][k]:e0["off"][k + 1]])
            assert p0["info"][slot, 0] == e0["info"][k, 0]
            k += 1
    assert k == 5


def test_shard_games_partition():
    from mlp_ppo_2ply_multi_b200.distributed import shard_games

    for total, world in ((65536, 8), (10, 4), (7, 8), (1, 1)):
        parts = [shard_games(total, r, world) for r in range(world)]
        assert sum(n for n, _ in parts) == total
        pos = 0
        for n, b in parts:
            assert b == pos
            pos += n


def test_temperature_schedule_matches_reference_formula():
    import mlp_ppo_2ply_multi_b200 as bg

    # reference src/multi/parameter_manager.py:93-111 with INITIAL 1.5, FINAL 0.5, MAX_UPDATES 4000
    assert bg.temperature_for_version(0) == 1.5 and bg.temperature_for_version(1) == 1.5
    assert bg.temperature_for_version(2001) == pytest.approx(1.0)
    assert bg.temperature_for_version(4001) == 0.5 and bg.temperature_for_version(10**6) == 0.5


def test_episode_record_semantics():
    import mlp_ppo_2ply_multi_b200 as bg

    ep = bg.Episode()
    e = bg.Experience(torch.zeros(198), 0.5, torch.tensor(0.3), False, torch.ones(198), 0.7)
    ep.add_experience(e, {"current_player": bg.Player.PLAYER1, "close_out_reward": True})
    ep.add_experience(bg.Experience(torch.zeros(198), 0.1, torch.tensor(1.0), True, torch.ones(198), 0.2),
                      {"current_player": bg.Player.PLAYER2, "win_type": "gammon", "winner": bg.Player.PLAYER2})
    assert ep.win_type == "gammon" and ep.close_out_counts == {bg.Player.PLAYER1: 1, bg.Player.PLAYER2: 0}
    assert ep.prime_reward_counts == {bg.Player.PLAYER1: 0, bg.Player.PLAYER2: 0}
    ep.to_numpy()
    assert isinstance(ep.experiences[0].observation, np.ndarray) and isinstance(ep.experiences[0].state_value, float)
    ep.to_tensor()
    x = ep.experiences[1]
    assert x.observation.dtype == torch.float32 and x.state_value.dtype == torch.float32 and x.done.dtype == torch.int64  # bool -> int64 as in the reference


def test_policy_network_state_dict_contract():
    import mlp_ppo_2ply_multi_b200 as bg

    net = bg.BackgammonPolicyNetwork()
    sd = net.state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == {"fc1.weight": (128, 198), "fc1.bias": (128,), "value_head.weight": (1, 128),
                                                          "value_head.bias": (1,)}
    x = torch.rand(5, 198)
    want = (torch.sigmoid(x @ sd["fc1.weight"].t() + sd["fc1.bias"]) @ sd["value_head.weight"].t() + sd["value_head.bias"]).squeeze(-1)
    assert torch.allclose(net(x), want, atol=1e-6)
    packed = bg.pack_weights(sd)
    assert packed.numel() == 200 * 128 + 1
    back = bg.unpack_weights(packed, 128)
    assert all(torch.equal(back[k], sd[k]) for k in sd)


def test_parameter_manager_surface_and_checkpoint_io(tmp_path, monkeypatch):
    """ParameterManager: the reference's 4-method surface (parameter_manager.py:54-111) and .pth files in its checkpoint format."""
    import mlp_ppo_2ply_multi_b200 as bg

    monkeypatch.chdir(tmp_path)
    pm = bg.ParameterManager(hidden_size=128)
    assert pm.get_version() == 1 and pm.get_temperature() == 1.5
    sd = pm.get_parameters()
    assert {k: tuple(v.shape) for k, v in sd.items()} == {"fc1.weight": (128, 198), "fc1.bias": (128,), "value_head.weight": (1, 128),
                                                          "value_head.bias": (1,)}
    net = bg.BackgammonPolicyNetwork()
    pm.set_parameters(net.state_dict())
    assert pm.get_version() == 2 and pm.get_temperature() == pytest.approx(1.5 - 1.0 / 4000)
    assert all(torch.equal(pm.get_parameters()[k], v) for k, v in net.state_dict().items())
    path = pm.save_model("ckpt.pth")
    loaded = torch.load(path, map_location="cpu")  # a plain state_dict, loadable by the reference's BackgammonPolicyNetwork
    net2 = bg.BackgammonPolicyNetwork()
    net2.load_state_dict(loaded)
    pm2 = bg.ParameterManager(hidden_size=128)
    pm2.load_model("ckpt.pth")
    assert pm2.get_version() == 2 and all(torch.equal(pm2.get_parameters()[k], v) for k, v in net.state_dict().items())
    # packed fast path == state-dict path
    pm2.set_packed(bg.pack_weights(net2.state_dict()))
    assert pm2.get_version() == 3 and torch.equal(bg.pack_weights(pm2.get_parameters()), bg.pack_weights(net.state_dict()))
    with pytest.raises(NotImplementedError):
        pm.save_model("x.pth", to_s3=True)


def test_reference_checkpoints_load_unchanged(golden):
    """The reference's shipped checkpoints (src/play/*.pth) load into BackgammonPolicyNetwork unchanged; the packed blob of the
    2.1 M-episode one is the golden `packed` vector.  Needs /root/reference (build container only)."""
    import glob

    import mlp_ppo_2ply_multi_b200 as bg

    files = sorted(glob.glob("/root/reference/src/play/*.pth"))
    if not files:
        pytest.skip("reference checkpoints not present on this machine")
    for f in files:
        sd = torch.load(f, map_location="cpu")
        net = bg.BackgammonPolicyNetwork(hidden_size=sd["fc1.weight"].shape[0])
        net.load_state_dict(sd)
        assert bg.pack_weights(net.state_dict()).numel() == 200 * net.hidden_size + 1
    sd = torch.load([f for f in files if f.endswith("backgammon_256_standard_episode_2100000.pth")][0], map_location="cpu")
    assert np.array_equal(bg.pack_weights(sd).numpy(), golden("values")["packed"])


def test_all_gather_episodes_single_process():
    """world size 1 (no process group): both layouts return the local batch unchanged"""
    from mlp_ppo_2ply_multi_b200 import distributed as bgd
    from mlp_ppo_2ply_multi_b200.episode import EpisodeBatch

    lens = [3, 1, 4]
    E, N = len(lens), sum(lens)
    g = torch.Generator().manual_seed(1)
    eb = EpisodeBatch(E, N, torch.randint(0, 6, (N + 5, 52), generator=g, dtype=torch.int8), torch.randint(0, 32, (N + 5,), generator=g, dtype=torch.uint8),
                      torch.rand(N + 5, generator=g), torch.rand(N + 5, generator=g), torch.rand(N + 5, generator=g),
                      torch.randint(1, 50, (N + 5,), generator=g, dtype=torch.int16), torch.randint(0, 50, (N + 5,), generator=g, dtype=torch.int16),
                      torch.randint(1, 7, (N + 5, 2), generator=g, dtype=torch.uint8), torch.tensor([0] + list(np.cumsum(lens)) + [0, 0], dtype=torch.int64),
                      torch.full((E + 2, 12), 7, dtype=torch.int32))
    m = bgd.all_gather_episodes(eb, 4, 12)
    assert (m.n_episodes, m.n_experiences, m.ep_offsets.tolist()) == (3, 8, [0, 3, 4, 8])
    assert torch.equal(m.after_boards, eb.after_boards[:N]) and torch.equal(m.reward, eb.reward[:N]) and torch.equal(m.roll, eb.roll[:N])
    p = bgd.all_gather_episodes(eb, 4, 12, compact=False)
    assert p.n_episodes == 4 and p.ep_len.tolist() == [3, 1, 4, 0] and p.ep_offsets[:4].tolist() == [0, 3, 4, 8]
    assert torch.equal(p.after_boards[:N], eb.after_boards[:N]) and p.episode_lengths().tolist() == [3, 1, 4, 0]
