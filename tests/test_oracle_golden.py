"""CPU tests: the C oracle (oracle/bg_oracle.c) against golden vectors produced by the unmodified Python
reference (tests/golden/make_golden.py).  This pins the oracle; the GPU tests then check CUDA == oracle."""
import numpy as np
import pytest


def test_opening_counts_appendix_b(oracle):
    # len(get_all_possible_moves(P, initial_board(), r)) over two_ply.DICE_ROLLS
    want = [42, 15, 16, 14, 8, 10, 75, 17, 18, 8, 14, 73, 17, 9, 14, 52, 9, 14, 4, 7, 11]
    b = oracle.initial_board()
    for p in (0, 1):
        assert [len(oracle.legal_moves(b, p, r)[0]) for r in oracle.DICE_ROLLS] == want


def test_quirk_q1(oracle):
    # appendix B: exactly one move (BAR->5), for both dice orders
    q = np.zeros(52, np.int8)
    q[10], q[48], q[50] = 1, 1, 13
    q[24 + 6], q[24 + 11], q[24 + 23] = 2, 2, 11
    for roll in ((6, 1), (1, 6)):
        boards, mv = oracle.legal_moves(q, 0, roll)
        assert len(boards) == 1
        assert tuple(mv[0, 0]) == (24, 5, 0) and mv[0, 1, 0] == 255


def test_movegen_golden_bit_exact(oracle, golden):
    g = golden("movegen")
    off, ob, om = oracle.movegen_batch(g["boards"], g["players"], g["rolls"])
    assert np.array_equal(off, g["offsets"])
    assert np.array_equal(ob, g["out_boards"])  # resulting boards, reference order
    assert np.array_equal(om, g["out_submoves"])  # sub-move sequences (start, end, hit)


def test_movegen_roll_order_irrelevant(oracle, golden):
    g = golden("movegen")
    sel = slice(0, 600)
    a = oracle.movegen_batch(g["boards"][sel], g["players"][sel], g["rolls"][sel])
    b = oracle.movegen_batch(g["boards"][sel], g["players"][sel], g["rolls"][sel][:, ::-1])
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_features_golden_bit_exact(oracle, golden):
    g = golden("features")
    f = oracle.encode(g["boards"], g["flags"])
    assert f.dtype == np.float32
    assert np.array_equal(f.view(np.uint32), g["features"].view(np.uint32))


def test_feature_lut_n_over_15(oracle):
    # appendix B: fp32 n/15 table
    want = "00000000 8988883d 8988083e cdcc4c3e 8988883e abaaaa3e cdcccc3e efeeee3e 8988083f 9a99193f abaa2a3f bcbb3b3f cdcc4c3f dedd5d3f efee6e3f 0000803f".split()
    b = np.zeros((16, 52), np.int8)
    b[:, 50] = np.arange(16)
    f = oracle.encode(b, np.zeros(16, np.uint8))
    got = [f[i, 193].tobytes().hex() for i in range(16)]
    assert got == want


def test_values_golden(oracle, golden):
    g = golden("values")
    H = int(g["H"])
    v = oracle.value(g["packed"], H, g["boards"], g["flags"])
    assert np.abs(v - g["values"]).max() < 1e-5  # north_star tolerance (fp32 torch vs double-accumulated oracle)
    v0 = oracle.value(g["packed_init0"], H, g["boards"], g["flags"])
    assert np.abs(v0 - g["values_init0"]).max() < 1e-5
    # opening values, appendix B
    b = np.stack([oracle.initial_board()] * 2)
    vo = oracle.value(g["packed"], H, b, np.array([0, 1], np.uint8))
    assert abs(vo[0] - 0.8230705) < 1e-5 and abs(vo[1] - 0.8347071) < 1e-5


def test_predicates_golden(oracle, golden):
    import ctypes as C

    g = golden("predicates")
    L = oracle.lib()
    fns = [L.bgo_check_game_over, L.bgo_check_gammon, L.bgo_check_backgammon, L.bgo_made_five_prime, L.bgo_is_closed_out]
    boards = np.ascontiguousarray(g["boards"])
    got = np.zeros_like(g["pred"])
    for i in range(len(boards)):
        ptr = boards[i].ctypes.data_as(C.c_void_p)
        for pl in (0, 1):
            got[i, pl] = [int(bool(f(ptr, pl))) for f in fns]
    assert np.array_equal(got, g["pred"])


@pytest.mark.parametrize("which", ["env_random", "env_shaping"])
def test_env_games_golden(oracle, golden, which):
    """the reference env's own games (random actions; a point-making policy that collects the +0.30 close-out and +0.20 prime
    rewards, backgammon_env.py:195-218) replayed through the oracle on the recorded dice"""
    g = golden(which)
    if which == "env_shaping":
        assert ((g["info"] & 4) != 0).sum() >= 1 and ((g["info"] & 8) != 0).sum() >= 5
        assert {20, 30} <= set(np.round(g["reward"] * 100).astype(int).tolist())
    n_games = len(g["tape_off"]) - 1
    for k in range(n_games):
        tape = g["tape"][g["tape_off"][k]:g["tape_off"][k + 1]]
        env = oracle.Env(tape=tape)
        env.reset()
        assert env.player == g["start_player"][k]
        assert env.num_moves == g["start_nmoves"][k]
        assert env.roll == tuple(g["start_roll"][k])
        for s in range(g["step_off"][k], g["step_off"][k + 1]):
            a = int(g["action"][s])
            r, done, info = env.step(None if a < 0 else a)
            assert r == pytest.approx(float(g["reward"][s]), abs=0) or abs(r - float(g["reward"][s])) < 1e-7
            assert int(done) == int(g["done"][s])
            assert (info & 0x1F) == (int(g["info"][s]) & 0x1F)
            assert np.array_equal(env.board, g["board"][s])
            assert env.player == g["player"][s]
            if not done:
                assert env.num_moves == g["nmoves"][s]
                assert env.roll == tuple(g["roll"][s])
            else:
                assert env.win_type == (int(g["info"][s]) >> 8)
        assert env.tape_pos == len(tape)  # consumed exactly the reference's dice, rejected doubles included


def test_env_truncate_golden(oracle, golden):
    """positions with more than 500 legal moves: the reference env keeps the first 500 (backgammon_env.py:262-272)"""
    g = golden("env_truncate")
    assert (g["true_count"] > 500).all()
    for i in range(len(g["boards"])):
        ob, _ = oracle.legal_moves(g["boards"][i], int(g["players"][i]), tuple(g["rolls"][i]))
        assert len(ob) == g["true_count"][i]
        assert np.array_equal(ob[:500], g["kept"][g["kept_off"][i]:g["kept_off"][i + 1]])


def test_worker_cap_golden(oracle, golden):
    """a game the unmodified Worker.play_episode cut at MAX_TIMESTEPS = 300 (worker.py:101): the reference's sampled actions replayed
    on its dice; 300 env steps, 295 experiences, no winner; values of the observation and of the chosen afterstate"""
    g = golden("worker_cap")
    packed, H = g["packed"], int(g["H"])
    env = oracle.Env(tape=g["tape"])
    env.reset()
    e = 0
    for t in range(int(g["n_steps"])):
        a = int(g["action"][t])
        if a < 0:
            assert env.num_moves == 0
            r, done, info = env.step(None)
            assert r == 0.0 and not done
            continue
        obs, mover = env.board, env.player
        after = env.afterstates[a]
        v = oracle.value(packed, H, np.stack([obs, after]), np.array([mover, mover], np.uint8))
        assert abs(v[0] - g["state_value"][e]) < 1e-5 and abs(v[1] - g["next_state_value"][e]) < 1e-5
        r, done, info = env.step(a)
        assert abs(r - float(g["reward"][e])) < 1e-7 and int(done) == int(g["done"][e]) == 0
        e += 1
    assert e == int(g["n_experiences"]) == 295 and int(g["n_steps"]) == 300 and int(g["win_type"]) == 0
    assert np.array_equal(env.board, g["final_board"]) and env.player == int(g["final_player"])
    assert env.tape_pos == len(g["tape"])


def test_greedy_games_golden(oracle, golden):
    g = golden("greedy_games")
    v = golden("values")
    packed, H = v["packed"], int(v["H"])
    n_games = len(g["tape_off"]) - 1
    for k in range(n_games):
        tape = g["tape"][g["tape_off"][k]:g["tape_off"][k + 1]]
        env = oracle.Env(tape=tape)
        stats, tr = env.play_episode(packed, H, temperature=0.0)
        lo, hi = g["dec_off"][k], g["dec_off"][k + 1]
        # near-ties (best-vs-second gap ~1e-6) may legitimately flip under a different fp32 summation order;
        # the golden games' smallest gap is 8e-6, well above fp32 noise (4e-7), so exact agreement is required
        assert g["gap"][lo:hi].min() > 2e-6
        assert stats["n_steps"] == g["n_steps"][k]
        assert stats["n_passes"] == g["n_passes"][k]
        assert stats["win_type"] == g["win_type"][k]
        assert stats["winner"] == g["winner"][k]
        assert np.array_equal(tr["nmoves"], g["nmoves"][lo:hi])
        assert np.array_equal(tr["action"], g["action"][lo:hi])  # move for move
        assert np.array_equal(tr["roll"], g["roll"][lo:hi])
        assert np.array_equal(tr["player"], g["player"][lo:hi])
        assert np.array_equal(tr["after"], g["after"][lo:hi])
        assert np.abs(tr["v"] - g["v"][lo:hi]).max() < 1e-5
        assert np.abs(tr["vnext"] - g["vnext"][lo:hi]).max() < 1e-5
        assert np.abs(tr["reward"] - g["reward"][lo:hi]).max() < 1e-7


def test_two_ply_golden(oracle, golden):
    g = golden("two_ply")
    v = golden("values")
    score, nrep = oracle.two_ply(g["cand_boards"], g["mover"], g["S"], v["packed"], int(v["H"]), top_k=5, alpha=1.0, beta=0.9)
    assert np.abs(score - g["score"]).max() < 1e-5
    assert (nrep > 0).all()
