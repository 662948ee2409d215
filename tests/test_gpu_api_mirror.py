"""GPU tests of the reference-facing Python surface: the same names, argument meaning, result order and
error behaviour as the reference, checked against vectors the unmodified reference produced (tests/golden/*.npz)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bg():
    import mlp_ppo_2ply_multi_b200 as m

    assert torch.cuda.is_available()
    return m


def test_get_all_possible_moves_and_execute_full_move(bg, golden):
    """get_all_possible_moves(player, board, roll_result) -> ordered List[FullMove] (generate_all_moves.py:7-66);
    execute_full_move_on_board_copy(board, full_move) -> ImmutableBoard (env_helper.py:27-91)."""
    g = golden("movegen")
    rng = np.random.default_rng(3)
    for i in rng.choice(len(g["players"]), 120, replace=False):
        board = bg.ImmutableBoard.from_array(g["boards"][i])
        player = bg.Player(int(g["players"][i]))
        moves = bg.get_all_possible_moves(player, board, [int(g["rolls"][i][0]), int(g["rolls"][i][1])])
        lo, hi = g["offsets"][i], g["offsets"][i + 1]
        assert len(moves) == hi - lo
        for k, fm in enumerate(moves):
            assert fm.player == player
            want = [tuple(int(x) for x in s) for s in g["out_submoves"][lo + k] if s[0] != 255]
            assert [(int(s.start), int(s.end), int(s.hits_blot)) for s in fm.sub_move_commands] == want
            after = bg.execute_full_move_on_board_copy(board, fm)
            assert np.array_equal(after.to_array(), g["out_boards"][lo + k])
            # chaining move_checker gives the same board (immutable_board.py:183-258)
            b2 = board
            for s in fm.sub_move_commands:
                b2 = b2.move_checker(player, s)
            assert b2 == after and hash(b2) == hash(after)


def test_board_features_and_generate_all_board_features(bg, golden):
    f = golden("features")
    for i in range(0, 1500, 50):
        b = bg.ImmutableBoard.from_array(f["boards"][i])
        x = b.get_board_features(bg.Player(int(f["flags"][i])))
        assert x.dtype == torch.float32 and x.shape == (198,) and x.device.type == "cpu"  # reference default device
        assert np.array_equal(x.numpy(), f["features"][i])
    g = golden("movegen")
    i = int(np.argmax(np.diff(g["offsets"])))  # the item with the most legal moves
    board, player = bg.ImmutableBoard.from_array(g["boards"][i]), bg.Player(int(g["players"][i]))
    moves = bg.get_all_possible_moves(player, board, list(map(int, g["rolls"][i])))
    X = bg.generate_all_board_features(board, player, moves)
    assert X.shape == (len(moves), 198)
    for k in (0, len(moves) // 2, len(moves) - 1):  # afterstates carry the MOVER's flag (env_helper.py:21)
        assert torch.equal(X[k], bg.execute_full_move_on_board_copy(board, moves[k]).get_board_features(player))
    assert bg.generate_all_board_features(board, player, []).shape == (0, 198)


class TapeEnv:
    """bg.BackgammonEnv whose dice come from a recorded tape (the golden games recorded np.random draws)"""

    def __new__(cls, bg, tape):
        class _E(bg.BackgammonEnv):
            def roll_dice(self_inner):
                self_inner.roll_result = [int(tape[self_inner._pos][0]), int(tape[self_inner._pos][1])]
                self_inner._pos += 1

        e = _E.__new__(_E)
        e._pos = 0
        _E.__init__(e)
        return e


@pytest.mark.parametrize("which", ["env_random", "env_shaping"])
def test_backgammon_env_replays_reference_games(bg, golden, capsys, which):
    """BackgammonEnv.reset/step against games of the unmodified reference env -- 12 under random actions (wins 1 / 2 / 2.5, passes) and
    40 under a point-making policy whose steps pay the +0.30 close-out and +0.20 prime rewards (tests/golden/make_golden_env_paths.py):
    boards, players, rolls, legal-move counts, rewards, done, info keys, observation layout."""
    g = golden(which)
    if which == "env_shaping":
        assert ((g["info"] & 4) != 0).any() and ((g["info"] & 8) != 0).any()
    games = range(len(g["tape_off"]) - 1)
    if which == "env_shaping":  # the env mirror is one launch chain per ply: replay the games that carry a shaping reward
        games = [k for k in games if ((g["info"][g["step_off"][k]:g["step_off"][k + 1]] & 12) != 0).any()][:8]
    for k in games:
        tape = g["tape"][g["tape_off"][k]:g["tape_off"][k + 1]]
        env = TapeEnv(bg, tape)
        obs = env.reset()
        assert int(env.current_player) == g["start_player"][k] and env.num_moves == g["start_nmoves"][k]
        assert tuple(env.roll_result) == tuple(g["start_roll"][k])
        assert obs.shape == (198,) and obs[196 + int(env.current_player)] == 1.0  # observations carry the player TO MOVE
        assert env.action_mask.sum().item() == env.num_moves and len(env.legal_moves) == env.num_moves
        for s in range(g["step_off"][k], g["step_off"][k + 1]):
            a = int(g["action"][s])
            mover = env.current_player
            obs, r, done, info = env.step(None if a < 0 else a)
            want_bits = int(g["info"][s])
            assert float(r) == pytest.approx(float(g["reward"][s]), abs=1e-7) and r.dim() == 0
            assert bool(done) == bool(g["done"][s]) and info["current_player"] == mover
            assert ("No legal" in str(info.get("info", ""))) == bool(want_bits & 1)
            assert bool(info.get("close_out_reward")) == bool(want_bits & 4) and bool(info.get("prime_reward")) == bool(want_bits & 8)
            assert np.array_equal(env.board.to_array(), g["board"][s]) and int(env.current_player) == g["player"][s]
            if done:
                assert {"regular": 1, "gammon": 2, "backgammon": 3}[info["win_type"]] == want_bits >> 8 and info["winner"] == mover
                assert env.game_over and env.win_type == info["win_type"]
            else:
                assert env.num_moves == g["nmoves"][s] and tuple(env.roll_result) == tuple(g["roll"][s])
                assert torch.equal(obs, env.board.get_board_features(env.current_player))
        assert env._pos == len(tape)  # consumed exactly the reference's dice, rejected doubles included
    # error behaviour: an invalid action is never raised, it costs -1 and changes nothing (backgammon_env.py:150-158)
    env = TapeEnv(bg, g["tape"][: g["tape_off"][1]])
    env.reset()
    before = env.board
    _, r, done, info = env.step(env.num_moves + 3)
    assert float(r) == -1.0 and not done and info["info"] == "Invalid action" and env.board == before
    assert "Invalid action" in capsys.readouterr().out


def test_policy_network_values_match_forward(bg, golden):
    """BackgammonPolicyNetwork: reference state_dict in, forward() == the CUDA evaluator on boards (policy_network.py:36-70)."""
    v = golden("values")
    H = int(v["H"])
    net = bg.BackgammonPolicyNetwork(hidden_size=H)
    net.load_state_dict(bg.unpack_weights(torch.from_numpy(v["packed"]), H))
    boards, flags = torch.from_numpy(v["boards"]).cuda(), torch.from_numpy(v["flags"]).cuda()
    got = net.values(boards, flags).cpu()
    assert torch.allclose(got, torch.from_numpy(v["values"]), atol=1e-5, rtol=0)  # vs the reference's torch forward
    x = bg.encode(boards, flags).cpu()
    assert torch.allclose(net(x), got, atol=1e-5, rtol=0)


def test_agent_play_step_plays_the_reference_greedy_game(bg, golden):
    """agent_play_step / select_highest_value_action (play_versus_ai.py:165-195) driving BackgammonEnv reproduce the reference's own
    greedy games (same dice): action for action, V(observation) and V(chosen afterstate) within 1e-5."""
    g, v = golden("greedy_games"), golden("values")
    H = int(v["H"])
    net = bg.BackgammonPolicyNetwork(hidden_size=H)
    net.load_state_dict(bg.unpack_weights(torch.from_numpy(v["packed"]), H))
    for k in range(2):
        tape = g["tape"][g["tape_off"][k]:g["tape_off"][k + 1]]
        env = TapeEnv(bg, tape)
        obs = env.reset()
        lo, hi = g["dec_off"][k], g["dec_off"][k + 1]
        d, done, steps = lo, False, 0
        while not done and steps < 300:
            if env.num_moves == 0:
                obs, _, done, _ = env.step(None)
            else:
                a = bg.agent_play_step(net, env)
                assert a == int(g["action"][d]) and env.num_moves == int(g["nmoves"][d])
                with torch.no_grad():
                    vals = net(torch.cat([obs.unsqueeze(0), env.legal_board_features[: env.num_moves]]))
                assert abs(float(vals[0]) - float(g["v"][d])) < 1e-5 and abs(float(vals[1 + a]) - float(g["vnext"][d])) < 1e-5
                obs, r, done, _ = env.step(a)
                assert float(r) == pytest.approx(float(g["reward"][d]), abs=1e-7)
                d += 1
            steps += 1
        assert d == hi and steps == int(g["n_steps"][k])
