"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs and against the golden vectors produced by the unmodified reference.  Bit-exact for boards / order /
sub-moves / features; |dV| <= 1e-5 for values (north_star contract)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def bg():
    import mlp_ppo_2ply_multi_b200 as m

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return m


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def run_movegen(bg, boards, players, rolls, **kw):
    res = bg.movegen(dev(boards), dev(players), dev(rolls), want_submoves=True, **kw)
    off, ob, om = res.canonical()
    return res, off.cpu().numpy(), ob.cpu().numpy(), om.cpu().numpy()


def test_movegen_golden_reference_vectors(bg, golden):
    g = golden("movegen")
    res, off, ob, om = run_movegen(bg, g["boards"], g["players"], g["rolls"], item_cap=4096)
    assert np.array_equal(res.counts.cpu().numpy(), np.diff(g["offsets"]))
    assert np.array_equal(off, g["offsets"])
    assert np.array_equal(ob, g["out_boards"])
    assert np.array_equal(om, g["out_submoves"])


def test_movegen_owner_and_offsets_consistent(bg, golden):
    g = golden("movegen")
    res = bg.movegen(dev(g["boards"]), dev(g["players"]), dev(g["rolls"]), item_cap=4096)
    total = res.total
    assert total == int(g["offsets"][-1])
    owner = res.owner[:total].cpu().numpy() if False else None
    offs, cnt = res.offsets.cpu().numpy(), res.counts.cpu().numpy()
    own = res.owner.cpu().numpy()
    # segments tile [0,total) exactly once and owner[] agrees
    order = np.argsort(offs[cnt > 0])
    starts = offs[cnt > 0][order]
    lens = cnt[cnt > 0][order]
    assert starts[0] == 0 and np.array_equal(starts[1:], (starts + lens)[:-1]) and starts[-1] + lens[-1] == total
    items = np.nonzero(cnt > 0)[0][order]
    assert np.array_equal(own[:total], np.repeat(items, lens))


@pytest.mark.parametrize("n_pos,seed", [(3000, 1), (20000, 2026)])
def test_movegen_vs_oracle_random_positions_all_rolls(bg, oracle, n_pos, seed):
    boards, players = oracle.random_positions(n_pos, seed=seed)
    ib, ip, ir = oracle.all_rolls_items(boards, players)
    o_off, o_b, o_m = oracle.movegen_batch(ib, ip, ir)
    res, off, ob, om = run_movegen(bg, ib, ip, ir, item_cap=4096, pool_cap=int(o_off[-1]) + 1024)
    assert np.array_equal(off, o_off)  # counts
    assert np.array_equal(ob, o_b)  # boards, reference order
    assert np.array_equal(om, o_m)  # sub-move sequences
    assert int(res.status_dev.item()) == 0


def test_movegen_roll_order_and_truncation(bg, oracle):
    boards, players = oracle.random_positions(2000, seed=5)
    ib, ip, ir = oracle.all_rolls_items(boards, players)
    a = run_movegen(bg, ib, ip, ir, item_cap=4096)
    b = run_movegen(bg, ib, ip, ir[:, ::-1].copy(), item_cap=4096)
    for x, y in zip(a[1:], b[1:]):
        assert np.array_equal(x, y)
    # item_cap keeps the FIRST entries (reference backgammon_env.py:262-272) while counts stay true
    cap = 7
    res, off, ob, om = run_movegen(bg, ib, ip, ir, item_cap=cap)
    true_cnt = np.diff(a[1])
    assert np.array_equal(res.counts.cpu().numpy(), true_cnt)
    keep = np.minimum(true_cnt, cap)
    assert np.array_equal(np.diff(off), keep)
    sel = np.concatenate([np.arange(s, s + k) for s, k in zip(a[1][:-1], keep)])
    assert np.array_equal(ob, a[2][sel])


def test_movegen_edge_cases(bg, oracle):
    # empty batch
    res = bg.movegen(torch.zeros((0, 52), dtype=torch.int8, device=DEV), torch.zeros(0, dtype=torch.uint8, device=DEV),
                     torch.zeros((0, 2), dtype=torch.uint8, device=DEV))
    assert res.total == 0
    # finished game (mover has 15 off) and fully blocked bar entry: zero moves
    z = np.zeros((2, 52), np.int8)
    z[0, 50] = 15
    z[0, 24 + 3] = 15
    z[1, 48] = 1
    z[1, 10] = 14
    z[1, 24:30] = 2
    z[1, 24 + 12] = 3
    players = np.zeros(2, np.uint8)
    rolls = np.array([[3, 4], [6, 6]], np.uint8)
    res, off, ob, om = run_movegen(bg, z, players, rolls)
    assert res.counts.cpu().tolist() == [0, 0]
    o_off, _, _ = oracle.movegen_batch(z, players, rolls)
    assert np.array_equal(off, o_off)
    # pool too small -> BG_ERR_CAPACITY reported, not silently truncated
    b = np.stack([oracle.initial_board()] * 8)
    with pytest.raises(bg.BgError) as ei:
        bg.movegen(dev(b), dev(np.zeros(8, np.uint8)), dev(np.array([[1, 1]] * 8, np.uint8)), pool_cap=100)
    assert ei.value.status == -3
    # invalid board (count > 15) -> BG_ERR_INVARIANT
    bad = oracle.initial_board().copy()
    bad[0] = 40
    with pytest.raises(bg.BgError) as ei:
        bg.movegen(dev(bad[None]), dev(np.zeros(1, np.uint8)), dev(np.array([[1, 2]], np.uint8)))
    assert ei.value.status == -4


def test_movegen_capacity_tiers(bg, oracle):
    """positions engineered to exceed the 128- and 1024-node tiers: many distinct points x small doubles"""
    rng = np.random.default_rng(3)
    boards = []
    for k in range(96):
        npts = [8, 10, 12, 13, 15, 15][k % 6]
        b = np.zeros(52, np.int8)
        pts = rng.choice(np.arange(0, 20), size=npts, replace=False)
        for p in pts:
            b[p] += 1
        b[int(pts[0])] += 15 - b[:24].sum()
        b[24 + 23] = 15  # opponent far away, nothing blocked
        boards.append(b)
    boards = np.array(boards, np.int8)
    players = np.zeros(len(boards), np.uint8)
    rolls = np.array([[1, 1]] * len(boards), np.uint8)
    o_off, o_b, o_m = oracle.movegen_batch(boards, players, rolls)
    cnt = np.diff(o_off)
    assert cnt.max() > 1024 and ((cnt > 128) & (cnt <= 1024)).any()  # exercises the 1024- and 4096-node tiers
    res, off, ob, om = run_movegen(bg, boards, players, rolls, item_cap=4096, pool_cap=int(o_off[-1]) + 16)
    assert np.array_equal(off, o_off) and np.array_equal(ob, o_b) and np.array_equal(om, o_m)


def test_truncation_at_500_matches_reference_golden(bg, golden):
    """positions with more than 500 legal moves: the reference env keeps the first 500 (backgammon_env.py:35,262-272); out_count is the true count"""
    g = golden("env_truncate")
    res = bg.movegen(dev(g["boards"]), dev(g["players"]), dev(g["rolls"]), item_cap=500, pool_cap=4096)
    assert np.array_equal(res.counts.cpu().numpy(), g["true_count"])
    off, ob, _ = res.canonical()
    assert np.array_equal(off.cpu().numpy(), g["kept_off"]) and np.array_equal(ob.cpu().numpy(), g["kept"])
    # the env mirror on the same positions
    for i in range(len(g["boards"])):
        env = bg.BackgammonEnv()
        env.reset()
        env.set_board(bg.ImmutableBoard.from_array(g["boards"][i]))
        env.set_current_player(bg.Player(int(g["players"][i])))
        env.roll_result = [int(x) for x in g["rolls"][i]]
        env.update_legal_moves()
        assert env.num_moves == 500 and len(env.legal_moves) == 500 and int(env.action_mask.sum().item()) == 500
        kept = g["kept"][g["kept_off"][i]:g["kept_off"][i + 1]]
        assert np.array_equal(np.asarray(env._legal_boards[:500]), kept)


def all_rolls_reference(oracle, boards, players, item_cap):
    ib, ip, ir = oracle.all_rolls_items(boards, players)
    off, ob, _ = oracle.movegen_batch(ib, ip, ir, want_moves=False)
    cnt = np.diff(off)
    if (cnt > item_cap).any():  # keep the first item_cap rows of every item
        keep = np.concatenate([np.arange(off[i], off[i] + min(cnt[i], item_cap)) for i in range(len(cnt))])
        ob = ob[keep]
        off = np.concatenate([[0], np.cumsum(np.minimum(cnt, item_cap))])
    return cnt, off, ob


def test_movegen_all_rolls_vs_oracle(bg, oracle, golden):
    """bg_movegen_all_rolls (one warp per position, all 21 rolls; csrc/movegen21.cu) == get_all_possible_moves per (position, roll):
    the reference's own golden positions, 6,000 random positions, the capacity-tier boards (doubles trees too wide for the fast path are
    handed to the per-item tiers), truncation at item_cap, inactive / invalid positions"""
    g = golden("movegen")
    # the golden file holds 21 consecutive items per position, in DICE_ROLLS order
    gb, gp = g["boards"][::21], g["players"][::21]
    assert np.array_equal(np.repeat(gb, 21, 0), g["boards"]) and np.array_equal(g["rolls"][:21], np.array(bg.DICE_ROLLS, np.uint8))
    res = bg.movegen_all_rolls(dev(gb), dev(gp), item_cap=4096)
    off, ob, _ = res.canonical()
    assert np.array_equal(off.cpu().numpy(), g["offsets"]) and np.array_equal(ob.cpu().numpy(), g["out_boards"])  # vs the reference itself
    res = bg.movegen_all_rolls(dev(gb), dev(gp), item_cap=4096, want_submoves=True)  # sub-moves route through the per-item kernels
    off, ob, om = res.canonical()
    assert np.array_equal(ob.cpu().numpy(), g["out_boards"]) and np.array_equal(om.cpu().numpy(), g["out_submoves"])
    boards, players = oracle.random_positions(6000, seed=77)
    # wide doubles trees (> the fast path's arena), bar / bear-off specials
    rng = np.random.default_rng(3)
    wide = []
    for k in range(48):
        npts = [8, 10, 12, 13, 15, 15][k % 6]
        b = np.zeros(52, np.int8)
        pts = rng.choice(np.arange(0, 20), size=npts, replace=False)
        for p in pts:
            b[p] += 1
        b[int(pts[0])] += 15 - b[:24].sum()
        b[24 + 23] = 15
        wide.append(b)
    boards = np.concatenate([boards, np.array(wide, np.int8)])
    players = np.concatenate([players, np.zeros(len(wide), np.uint8)])
    for cap in (4096, 500, 40):
        cnt, off, ob = all_rolls_reference(oracle, boards, players, cap)
        res = bg.movegen_all_rolls(dev(boards), dev(players), item_cap=cap, pool_cap=int(off[-1]) + 64, want_owner=True)
        assert np.array_equal(res.counts.cpu().numpy(), cnt)
        assert res.total == int(off[-1])
        o2, b2, _ = res.canonical()
        assert np.array_equal(o2.cpu().numpy(), off) and np.array_equal(b2.cpu().numpy(), ob)
        T = res.total
        own = res.owner[:T].to(torch.int64)
        assert bool((res.flags[:T] == dev(players)[own // 21]).all())
    assert cnt.max() > 1024
    # a pool that is too small: BG_ERR_CAPACITY, the items that did not fit are marked, nothing is written out of bounds
    small = int(off[-1]) // 3
    guard = torch.full((small + 4096, 52), 77, dtype=torch.int8, device="cuda")
    res = bg.movegen_all_rolls(dev(boards), dev(players), item_cap=40, pool_cap=small, out_boards=guard[:small], check_status=False)
    assert int(res.status_dev.item()) == -3 and bool((guard[small:] == 77).all())
    o = res.offsets.cpu().numpy()
    assert (o == -1).any() and np.array_equal(res.counts.cpu().numpy(), cnt)
    ok_items = np.nonzero(o >= 0)[0]
    assert ((o[ok_items] + np.minimum(cnt[ok_items], 40)) <= small).all()
    # an invalid board poisons only its own position
    bad = boards[:64].copy()
    bad[5, 3] = 17
    res = bg.movegen_all_rolls(dev(bad), dev(players[:64]), item_cap=4096, check_status=False)
    assert int(res.status_dev.item()) == -4
    c = res.counts.cpu().numpy().reshape(64, 21)
    cnt64 = cnt[:64 * 21].reshape(64, 21)
    assert (c[5] == 0).all() and np.array_equal(np.delete(c, 5, 0), np.delete(cnt64, 5, 0))


def test_movegen_eval_all_rolls_equals_item_form(bg, oracle, golden):
    """bg_movegen_eval_all_rolls == bg_movegen_eval on the replicated items: counts, boards, values (<= 1e-5 vs the double oracle)"""
    v = golden("values")
    H = int(v["H"])
    w = bg.prepare_weights(dev(v["packed"]), H)
    boards, players = oracle.random_positions(3000, seed=5)
    cnt, off, ob = all_rolls_reference(oracle, boards, players, 500)
    cap = int(off[-1]) + 4096
    pool = torch.empty((cap, 52), dtype=torch.int8, device="cuda")
    flags = torch.empty(cap, dtype=torch.uint8, device="cuda")
    vals = torch.empty(cap, dtype=torch.float32, device="cuda")
    res, vals = bg.movegen_evaluate_all_rolls(dev(boards), dev(players), w, pool, flags, vals, item_cap=500, check_status=True)
    assert np.array_equal(res.counts.cpu().numpy(), cnt) and res.total == int(off[-1])
    o2, b2, _ = res.canonical()
    assert np.array_equal(b2.cpu().numpy(), ob)
    kept = torch.clamp(res.counts.to(torch.int64), max=500)
    item = torch.repeat_interleave(torch.arange(kept.numel(), device="cuda"), kept)
    rows = res.offsets[item] + (torch.arange(item.numel(), device="cuda") - o2[item])
    want = oracle.value(v["packed"], H, ob, np.repeat(np.repeat(players, 21), cnt.clip(max=500)))
    assert np.abs(vals[rows].cpu().numpy() - want).max() < 1e-5


def test_compact_pool_equals_board_pool(bg, oracle, golden):
    """bg_movegen[_eval]_all_rolls_compact (8-byte (code, position) rows, afterstates rebuilt inside the evaluator) == the board forms:
    counts, order, values; bg_afterstates_from_codes rebuilds every board bit-exactly (vs the oracle), also under item_cap truncation"""
    v = golden("values")
    H = int(v["H"])
    w = bg.prepare_weights(dev(v["packed"]), H)
    boards, players = oracle.random_positions(4000, seed=123)
    rng = np.random.default_rng(5)
    wide = []
    for k in range(24):  # wide doubles trees: the big code tier
        b = np.zeros(52, np.int8)
        pts = rng.choice(np.arange(0, 20), size=[8, 10, 12, 13, 15, 15][k % 6], replace=False)
        for p in pts:
            b[p] += 1
        b[int(pts[0])] += 15 - b[:24].sum()
        b[24 + 23] = 15
        wide.append(b)
    boards = np.concatenate([boards, np.array(wide, np.int8)])
    players = np.concatenate([players, np.zeros(24, np.uint8)])
    for cap in (4096, 500, 40):
        cnt, off, ob = all_rolls_reference(oracle, boards, players, cap)
        T = int(off[-1])
        res, vals = bg.movegen_all_rolls_compact(dev(boards), dev(players), w, item_cap=cap, pool_cap=T + 4096, check_status=True)
        assert np.array_equal(res.counts.cpu().numpy(), cnt) and res.total == T
        kept = torch.clamp(res.counts.to(torch.int64), max=cap)
        o2 = torch.zeros(kept.numel() + 1, dtype=torch.int64, device="cuda")
        o2[1:] = torch.cumsum(kept, 0)
        item = torch.repeat_interleave(torch.arange(kept.numel(), device="cuda"), kept)
        rows = res.offsets[item] + (torch.arange(T, device="cuda") - o2[item])
        assert np.array_equal(res.afterstates(rows).cpu().numpy(), ob)  # every afterstate, in action order, vs the oracle
        want = oracle.value(v["packed"], H, ob, np.repeat(np.repeat(players, 21), np.minimum(cnt, cap)))
        assert np.abs(vals[rows].cpu().numpy() - want).max() < 1e-5
        v2 = bg.evaluate_codes(res, w)
        assert torch.equal(v2[rows], vals[rows])
        if cap == 500:  # same tiles, same operands: the board path gives the same bits
            pool = torch.empty((T + 4096, 52), dtype=torch.int8, device="cuda")
            flags = torch.empty(T + 4096, dtype=torch.uint8, device="cuda")
            vb = torch.empty(T + 4096, dtype=torch.float32, device="cuda")
            rb, vb = bg.movegen_evaluate_all_rolls(dev(boards), dev(players), w, pool, flags, vb, item_cap=cap)
            kb = torch.clamp(rb.counts.to(torch.int64), max=cap)
            rows_b = rb.offsets[item] + (torch.arange(T, device="cuda") - o2[item])
            assert torch.equal(kb, kept) and (vb[rows_b] - vals[rows]).abs().max().item() < 1e-6
            act = bg.select(vals, res.offsets, res.counts, temperature=0.0, item_cap=cap)
            chosen = res.chosen_afterstates(act).cpu().numpy()
            a = act.cpu().numpy()
            for i in np.random.default_rng(1).choice(len(a), 300, replace=False):
                if a[i] >= 0:
                    assert np.array_equal(chosen[i], ob[off[i] + a[i]])
                else:
                    assert cnt[i] == 0 and not chosen[i].any()


def test_encode_bit_exact(bg, oracle, golden):
    g = golden("features")
    f = bg.encode(dev(g["boards"]), dev(g["flags"])).cpu().numpy()
    assert np.array_equal(f.view(np.uint32), g["features"].view(np.uint32))  # vs the reference itself
    boards, players = oracle.random_positions(20000, seed=9)
    f = bg.encode(dev(boards), dev(players)).cpu().numpy()
    assert np.array_equal(f.view(np.uint32), oracle.encode(boards, players).view(np.uint32))


@pytest.mark.parametrize("which", ["packed", "packed_init0"])
def test_eval_values(bg, oracle, golden, which):
    g = golden("values")
    H = int(g["H"])
    w = bg.prepare_weights(dev(g[which]), H)
    v = bg.evaluate(dev(g["boards"]), dev(g["flags"]), w).cpu().numpy()
    ref = g["values" if which == "packed" else "values_init0"]
    assert np.abs(v - ref).max() < 1e-5  # vs torch reference forward
    boards, players = oracle.random_positions(5000, seed=4)
    ib, ip, ir = oracle.all_rolls_items(boards[:1500], players[:1500])
    o_off, o_b, _ = oracle.movegen_batch(ib, ip, ir, want_moves=False)
    flags = np.repeat(ip, np.diff(o_off))
    v = bg.evaluate(dev(o_b), dev(flags), w).cpu().numpy()
    v_ref = oracle.value(g[which], H, o_b, flags)  # double-accumulated oracle
    assert np.abs(v - v_ref).max() < 1e-5
    # owner indirection gives identical results
    owner = np.repeat(np.arange(len(ip), dtype=np.int32), np.diff(o_off))
    v2 = bg.evaluate(dev(o_b), None, w, owner=dev(owner), owner_players=dev(ip)).cpu().numpy()
    assert np.abs(v - v2).max() < 2e-6  # owner indirection runs the FFMA kernel, per-row flags (large batch) the tcgen05 kernel


@pytest.mark.parametrize("H", [32, 64, 96, 256])
def test_eval_other_hidden_sizes(bg, oracle, H):
    rng = np.random.default_rng(H)
    packed = (rng.standard_normal(200 * H + 1) * 0.5).astype(np.float32)
    boards, players = oracle.random_positions(3000, seed=H)
    w = bg.prepare_weights(dev(packed), H)
    v = bg.evaluate(dev(boards), dev(players), w).cpu().numpy()
    assert np.abs(v - oracle.value(packed, H, boards, players)).max() < 1e-5 * max(1.0, np.sqrt(H / 128))


def test_select_greedy_and_distribution(bg):
    rng = np.random.default_rng(0)
    counts = rng.integers(0, 60, size=4000).astype(np.int32)
    counts[:5] = [0, 1, 500, 33, 2]
    off = np.zeros(len(counts) + 1, np.int64)
    off[1:] = np.cumsum(counts)
    v = rng.standard_normal(off[-1]).astype(np.float32)
    v[off[3]:off[3] + 33] = 0.25  # exact ties -> lowest index
    act = bg.select(dev(v), dev(off[:-1]), dev(counts), temperature=0.0).cpu().numpy()
    want = np.array([-1 if c == 0 else int(np.argmax(v[o:o + c])) for o, c in zip(off[:-1], counts)])
    assert np.array_equal(act, want) and act[3] == 0
    # softmax(V/T) sampling: chi-square of empirical frequencies against the exact distribution
    n, T, trials = 12, 0.7, 40000
    vals = rng.standard_normal(n).astype(np.float32)
    p = np.exp(vals / T - (vals / T).max())
    p /= p.sum()
    vv = np.tile(vals, trials)
    offs = np.arange(trials, dtype=np.int64) * n
    cn = np.full(trials, n, np.int32)
    a = bg.select(dev(vv), dev(offs), dev(cn), temperature=T, seed=123, ctr=7).cpu().numpy()
    freq = np.bincount(a, minlength=n)
    chi2 = ((freq - trials * p) ** 2 / (trials * p)).sum()
    assert chi2 < 40.0  # dof = 11; P(chi2 > 40) ~ 4e-5
    a2 = bg.select(dev(vv), dev(offs), dev(cn), temperature=T, seed=123, ctr=7).cpu().numpy()
    assert np.array_equal(a, a2)  # counter-based RNG: reproducible
    a3 = bg.select(dev(vv), dev(offs), dev(cn), temperature=T, seed=123, ctr=8).cpu().numpy()
    assert not np.array_equal(a, a3)


def test_full_size_properties(bg, oracle):
    """BASELINE config-2 scale slice (131,072 positions x 21 rolls = 2.75 M items, 58 M afterstates): counts vs oracle, checker
    conservation, pip monotonicity, and an order-sensitive 64-bit checksum of every item's afterstate list vs the oracle."""
    n_pos = 131072
    boards, players = oracle.random_positions(n_pos, seed=2026)
    ib, ip, ir = oracle.all_rolls_items(boards, players)
    o_off, o_boards, _ = oracle.movegen_batch(ib, ip, ir, want_moves=False)
    total = int(o_off[-1])
    res = bg.movegen(dev(ib), dev(ip), dev(ir), item_cap=4096, pool_cap=total + 4096, want_owner=True)
    assert res.total == total
    assert np.array_equal(res.counts.cpu().numpy(), np.diff(o_off))
    ob = res.boards[:total].to(torch.int32)
    own = res.owner[:total].long()
    pl = dev(ip)[own].long()
    # 15 checkers per side on every afterstate
    s0 = ob[:, 0:24].sum(1) + ob[:, 48] + ob[:, 50]
    s1 = ob[:, 24:48].sum(1) + ob[:, 49] + ob[:, 51]
    assert bool((s0 == 15).all()) and bool((s1 == 15).all())
    # pip count of the mover never increases
    w0 = torch.arange(24, 0, -1, device=DEV, dtype=torch.int32)
    w1 = torch.arange(1, 25, device=DEV, dtype=torch.int32)
    root = dev(ib).to(torch.int32)[own]

    def pips(b):
        p0 = (b[:, 0:24] * w0).sum(1) + 25 * b[:, 48]
        p1 = (b[:, 24:48] * w1).sum(1) + 25 * b[:, 49]
        return torch.where(pl == 0, p0, p1)

    assert bool((pips(ob) < pips(root)).all())
    del ob, root
    # content AND order of all 58 M afterstates: a 64-bit checksum per item, sum_k (2k+1) * hash(board_k), vs the oracle
    # ("boards, order, counts ... via 64-bit checksum per item")
    coef = (np.arange(1, 14, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) | np.uint64(1)
    words = np.ascontiguousarray(o_boards).view(np.uint32).reshape(-1, 13)
    counts = np.diff(o_off)
    k = np.arange(total, dtype=np.int64) - np.repeat(o_off[:-1], counts)
    contrib = np.empty(total, np.uint64)
    CH = 1 << 22
    for lo in range(0, total, CH):
        hi = min(total, lo + CH)
        h = (words[lo:hi].astype(np.uint64) * coef).sum(1, dtype=np.uint64)
        h ^= h >> np.uint64(29)
        contrib[lo:hi] = h * (2 * k[lo:hi] + 1).astype(np.uint64)
    want = np.zeros(len(ib), np.uint64)
    nz = counts > 0
    want[nz] = np.add.reduceat(contrib, o_off[:-1][nz])  # CSR segments (empty items skipped: reduceat would repeat a neighbour)
    got = torch.zeros(len(ib), dtype=torch.int64, device=DEV)
    coef_t = torch.from_numpy(coef.view(np.int64)).to(DEV)
    gw = res.boards[:total].view(torch.int32).reshape(-1, 13)
    for lo in range(0, total, CH):
        hi = min(total, lo + CH)
        w64 = gw[lo:hi].to(torch.int64) & 0xFFFFFFFF
        h = (w64 * coef_t).sum(1)
        h = h ^ ((h >> 29) & ((1 << 35) - 1))  # logical shift
        rows = torch.arange(lo, hi, device=DEV)
        kk = rows - res.offsets[own[lo:hi]]
        got.index_add_(0, own[lo:hi], h * (2 * kk + 1))
    assert np.array_equal(got.cpu().numpy().view(np.uint64), want)


def test_two_ply_reference_setting_and_best_reply(bg, oracle, golden):
    """bg_two_ply vs compute_weighted_opponent_response of the reference (golden, top-5 / alpha 1 / beta 0.9) and vs the oracle"""
    g = golden("two_ply")
    v = golden("values")
    H = int(v["H"])
    w = bg.prepare_weights(dev(v["packed"]), H)
    score, nrep = bg.two_ply(dev(g["cand_boards"]), dev(g["mover"]), dev(g["S"]), w, top_k=5, alpha=1.0, beta=0.9)
    assert np.abs(score.cpu().numpy() - g["score"]).max() < 1e-5  # vs the reference itself
    # larger sample vs the oracle, both settings, small workspace to force chunking
    boards, players = oracle.random_positions(600, seed=31)
    ib, ip, ir = oracle.all_rolls_items(boards[:40], players[:40])
    o_off, o_b, _ = oracle.movegen_batch(ib, ip, ir, want_moves=False)
    mover = np.repeat(ip, np.diff(o_off))
    sel = np.random.default_rng(1).choice(len(o_b), size=700, replace=False)
    cb, mv = o_b[sel], mover[sel]
    S = oracle.value(v["packed"], H, cb, mv)
    ws_small = torch.empty(bg._lib.lib().bg_two_ply_workspace_bytes(64), dtype=torch.uint8, device=DEV)
    for k, a, b in ((5, 1.0, 0.9), (1, 1.0, 1.0)):
        want, want_rep = oracle.two_ply(cb, mv, S, v["packed"], H, top_k=k, alpha=a, beta=b)
        got, rep = bg.two_ply(dev(cb), dev(mv), dev(S), w, top_k=k, alpha=a, beta=b)
        assert np.abs(got.cpu().numpy() - want).max() < 1e-5
        assert np.array_equal(rep.cpu().numpy(), want_rep)
        got2, _ = bg.two_ply(dev(cb), dev(mv), dev(S), w, top_k=k, alpha=a, beta=b, workspace=ws_small)
        assert np.array_equal(got2.cpu().numpy(), got.cpu().numpy())


def _philox4x32(key, ctr):
    """Philox4x32-10 (csrc/bg_common.cuh Philox::gen with ctr_hi = 0)"""
    M0, M1, m32 = 0xD2511F53, 0xCD9E8D57, 0xFFFFFFFF
    k0, k1 = key & m32, (key >> 32) & m32
    c = [ctr & m32, (ctr >> 32) & m32, 0, 0]
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & m32, p1 & m32, ((p0 >> 32) ^ c[3] ^ k1) & m32, p0 & m32]
        k0, k1 = (k0 + 0x9E3779B9) & m32, (k1 + 0xBB67AE85) & m32
    return c


def _mix32(a, b, c, d):
    m32 = 0xFFFFFFFF
    h = (a * 0x9E3779B1) & m32
    h = ((h ^ (h >> 15)) + b * 0x85EBCA77) & m32
    h = ((h ^ (h >> 13)) + c * 0xC2B2AE3D) & m32
    h = ((h ^ (h >> 16)) + d * 0x27D4EB2F) & m32
    h ^= h >> 15
    h = (h * 0x2C1B3C6D) & m32
    return h ^ (h >> 12)


def _sample_rows(seed, item, n, cap):
    """rows perm(0) .. perm(cap - 1) of an item with n replies, as documented at bg_two_ply_reply_sampling (include/bgarena.h)"""
    k = _philox4x32(seed ^ 0x3C6EF372FE94F82B, item)
    bits = 2
    while (1 << bits) < n:
        bits += 2
    half, out = bits >> 1, []
    mask = (1 << half) - 1
    for j in range(cap):
        x = j
        while True:
            L, R = x >> half, x & mask
            for rd in range(4):
                L, R = R, L ^ (_mix32(R, k[rd], rd, 0x9E3779B9) & mask)
            x = (L << half) | R
            if x < n:
                break
        out.append(x)
    return out


def test_two_ply_reply_sampling_option(bg, oracle, golden):
    """bg_two_ply_reply_sampling: the reference's random.sample(opponent_moves, 50) on 1-1 / 2-2 / 3-3 (two_ply.py:119-121) as a reproducible
    option.  The sampled rows are recomputed here from the documented permutation (Philox keys, 4-round Feistel, cycle walking), the expected
    scores from the oracle's replies and values: exact reply counts, scores within 1e-5; the sample is a subset, so no score drops below the
    unsampled one (beta > 0); same seed -> same bits, other seed -> other sample; a cap above every count changes nothing; chunking (a
    small workspace) does not change the sample."""
    v = golden("values")
    H = int(v["H"])
    w = bg.prepare_weights(dev(v["packed"]), H)
    boards, players = oracle.random_positions(400, seed=61)
    ib, ip, ir = oracle.all_rolls_items(boards[:30], players[:30])
    o_off, o_b, _ = oracle.movegen_batch(ib, ip, ir, want_moves=False)
    mover = np.repeat(ip, np.diff(o_off))
    sel = np.random.default_rng(3).choice(len(o_b), size=160, replace=False)
    cb, mv = o_b[sel], mover[sel]
    S = oracle.value(v["packed"], H, cb, mv)
    cap, seed, top_k, alpha, beta = 50, 1234567, 5, 1.0, 0.9
    # the oracle's replies of every (candidate, roll) item, in action order, and their values
    rb, rp, rr = oracle.all_rolls_items(cb, 1 - mv)
    r_off, r_b, _ = oracle.movegen_batch(rb, rp, rr, want_moves=False)
    r_val = oracle.value(v["packed"], H, r_b, np.repeat(rp, np.diff(r_off)))
    want = np.zeros(len(cb))
    want_rep = np.zeros(len(cb), np.int64)
    n_sampled_items = 0
    for c in range(len(cb)):
        W = 0.0
        for r in range(21):
            it = c * 21 + r
            vals = r_val[r_off[it]:r_off[it + 1]]
            n = len(vals)
            if n == 0:
                continue
            if r in (0, 6, 11) and n > cap:
                vals = vals[_sample_rows(seed, it, n, cap)]
                n_sampled_items += 1
            want_rep[c] += len(vals)
            top = np.sort(vals)[::-1][:top_k]
            W += float(top.astype(np.float32).mean()) * ((1.0 if r in (0, 6, 11, 15, 18, 20) else 2.0) / 36.0)
        want[c] = alpha * S[c] - beta * W
    assert n_sampled_items > 50  # the option has something to do on this sample
    full, full_rep = bg.two_ply(dev(cb), dev(mv), dev(S), w, top_k=top_k, alpha=alpha, beta=beta)
    prev = bg.set_reply_sampling(cap, seed)
    try:
        assert prev == 0
        got, rep = bg.two_ply(dev(cb), dev(mv), dev(S), w, top_k=top_k, alpha=alpha, beta=beta)
        assert np.array_equal(rep.cpu().numpy(), want_rep)
        assert np.abs(got.cpu().numpy() - want).max() < 1e-5
        assert (got - full).min().item() > -1e-6 and (rep <= full_rep).all() and (rep < full_rep).any()
        again, _ = bg.two_ply(dev(cb), dev(mv), dev(S), w, top_k=top_k, alpha=alpha, beta=beta)
        assert torch.equal(again, got)
        ws_small = torch.empty(bg._lib.lib().bg_two_ply_workspace_bytes(24), dtype=torch.uint8, device=DEV)
        chunked, _ = bg.two_ply(dev(cb), dev(mv), dev(S), w, top_k=top_k, alpha=alpha, beta=beta, workspace=ws_small)
        assert torch.equal(chunked, got)
        bg.set_reply_sampling(cap, seed + 1)
        other, rep2 = bg.two_ply(dev(cb), dev(mv), dev(S), w, top_k=top_k, alpha=alpha, beta=beta)
        assert torch.equal(rep2, rep) and not torch.equal(other, got)
        bg.set_reply_sampling(100000, seed)
        same, rep3 = bg.two_ply(dev(cb), dev(mv), dev(S), w, top_k=top_k, alpha=alpha, beta=beta)
        assert torch.equal(same, full) and torch.equal(rep3, full_rep)
    finally:
        bg.set_reply_sampling(0, 0)
    off, _ = bg.two_ply(dev(cb), dev(mv), dev(S), w, top_k=top_k, alpha=alpha, beta=beta)
    assert torch.equal(off, full)


@pytest.mark.parametrize("which", ["packed", "packed_init0"])
def test_eval_tensor_core_and_ffma_kernels_agree_with_oracle(bg, oracle, golden, which):
    """H = 128 has two evaluators: tcgen05/TMEM (two fp16 weight pieces, batches >= 32768 rows) and FFMA gather (smaller
    batches).  Both must meet the 1e-5 contract against the double-accumulated oracle, including ragged tails and stacks > 6."""
    g = golden("values")
    w = bg.prepare_weights(dev(g[which]), 128)
    boards, players = oracle.random_positions(2500, seed=77)
    ib, ip, ir = oracle.all_rolls_items(boards, players)
    o_off, o_b, _ = oracle.movegen_batch(ib, ip, ir, want_moves=False)
    flags = np.repeat(ip, np.diff(o_off))
    extra = np.zeros((64, 52), np.int8)  # tall stacks, bar and borne-off checkers
    extra[:, 5] = np.arange(64) % 16
    extra[:, 24 + 18] = 15 - (np.arange(64) % 16)
    extra[:, 48] = np.arange(64) % 3
    extra[:, 50] = 15 - extra[:, 5] - extra[:, 48]
    extra[:, 51] = np.arange(64) % 16
    o_b = np.concatenate([o_b, extra])
    flags = np.concatenate([flags, (np.arange(64) % 2).astype(np.uint8)])
    n = (len(o_b) // 128) * 128 - 37  # ragged last tile
    o_b, flags = o_b[:n], flags[:n]
    assert n > 32768
    ref = oracle.value(g[which], 128, o_b, flags)
    v_tc = bg.evaluate(dev(o_b), dev(flags), w).cpu().numpy()
    assert bg._lib.lib().bg_eval_tc_status() == 0
    assert np.abs(v_tc - ref).max() < 1e-5
    v_ff = np.concatenate([bg.evaluate(dev(o_b[i:i + 20000]), dev(flags[i:i + 20000]), w).cpu().numpy() for i in range(0, n, 20000)])
    assert np.abs(v_ff - ref).max() < 1e-5
    assert np.abs(v_ff - v_tc).max() < 2e-6


@pytest.mark.parametrize("schedule", [0, 1])
def test_eval_tensor_core_tile_schedules(bg, oracle, golden, schedule):
    """The tcgen05 evaluator hands local tile n of a CTA to TMEM slot n & 1, builder set n & 1 and the epilogue warps of that slot: row counts
    around the schedule's edges -- one tile per CTA, an odd number of tiles (a CTA whose second slot is never used), exactly one / two tiles
    for each of the 148 CTAs, a ragged last tile -- must give the FFMA kernel's values and leave the status word clear, under the static
    split (0) and under the dynamic one (1: tile pairs claimed from a grid-wide counter), for board pools and for compact (code) pools."""
    g = golden("values")
    w = bg.prepare_weights(dev(g["packed"]), 128)
    boards, players = oracle.random_positions(3 * 37888 + 300, seed=4242)
    d_b, d_p = dev(boards), dev(players)
    v_ff = torch.cat([bg.evaluate(d_b[i:i + 30000], d_p[i:i + 30000], w) for i in range(0, len(boards), 30000)])
    ref = oracle.value(g["packed"], 128, boards[:4096], players[:4096])
    assert np.abs(v_ff[:4096].cpu().numpy() - ref).max() < 1e-5
    prev = bg._lib.lib().bg_eval_tc_tile_schedule(schedule)
    try:
        for n in (32768, 32769, 32768 + 3 * 128 + 5, 148 * 128 * 2 - 1, 148 * 128 * 2, 148 * 128 * 2 + 1, 148 * 128 * 3, 148 * 128 * 3 + 129, 3 * 37888 + 300):
            v_tc = bg.evaluate(d_b[:n], d_p[:n], w)
            assert bg._lib.lib().bg_eval_tc_status() == 0
            assert (v_tc - v_ff[:n]).abs().max().item() < 2e-6, n
        # compact pools always run on the tensor-core kernel: all 21 rolls of 1,500 positions, values against the FFMA kernel on the
        # afterstates materialised from the codes
        pb, pp = oracle.random_positions(1500, seed=99)
        res, vals = bg.movegen_all_rolls_compact(dev(pb), dev(pp), w, item_cap=500, check_status=True)
        assert bg._lib.lib().bg_eval_tc_status() == 0
        T = int(res.total)
        ab = res.afterstates(torch.arange(T, device="cuda"))
        fl = dev(pp)[(res.codes[:T] >> 32).to(torch.int64)]
        v_f = torch.cat([bg.evaluate(ab[i:i + 30000], fl[i:i + 30000], w) for i in range(0, T, 30000)])
        assert (vals[:T] - v_f).abs().max().item() < 2e-6
    finally:
        bg._lib.lib().bg_eval_tc_tile_schedule(prev)


@pytest.mark.parametrize("H", [32, 64, 96, 160, 224, 256])
def test_eval_other_nets_on_the_tensor_core_path(bg, oracle, H):
    """every hidden size runs on the tcgen05 kernel for batches >= 32768 rows: fewer than 128 units zero-padded to 128, more than 128 in
    two passes of 128 units (the second accumulates); the CUDA-core kernel serves smaller batches.  Both against the oracle (1e-5) and
    against each other."""
    rng = np.random.default_rng(1000 + H)
    packed = (rng.standard_normal(200 * H + 1) * 0.5).astype(np.float32)
    boards, players = oracle.random_positions(40000, seed=H)
    w = bg.prepare_weights(dev(packed), H)
    ref = oracle.value(packed, H, boards, players)
    v_tc = bg.evaluate(dev(boards), dev(players), w).cpu().numpy()  # 40,000 rows: tensor-core path
    assert bg._lib.lib().bg_eval_tc_status() == 0
    v_ff = np.concatenate([bg.evaluate(dev(boards[i:i + 20000]), dev(players[i:i + 20000]), w).cpu().numpy() for i in (0, 20000)])
    tol = 1e-5 * max(1.0, np.sqrt(H / 128))
    assert np.abs(v_tc - ref).max() < tol and np.abs(v_ff - ref).max() < tol and np.abs(v_tc - v_ff).max() < tol


def test_host_pipeline_equals_resident_path(bg, oracle):
    """bg.HostPipeline (pinned host batch, chunks on two streams) must give the same counts and greedy actions as one resident call."""
    boards, players = oracle.random_positions(6000, seed=31)
    ib, ip, ir = oracle.all_rolls_items(boards, players)
    B = ib.shape[0]
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "values.npz"))
    w = bg.prepare_weights(torch.from_numpy(g["packed"]).to(DEV), int(g["H"]))
    db, dp, dr = (torch.from_numpy(x).to(DEV) for x in (ib, ip, ir))
    res = bg.movegen(db, dp, dr, item_cap=500)
    v = bg.evaluate(res.boards, res.flags, w, n_dev=res.total_dev)
    act = bg.select(v, res.offsets, res.counts, temperature=0.0)
    hb, hp, hr = (torch.from_numpy(x).pin_memory() for x in (ib, ip, ir))
    ha, hc = torch.full((B,), -7, dtype=torch.int32).pin_memory(), torch.full((B,), -7, dtype=torch.int32).pin_memory()
    pipe = bg.HostPipeline(w, items_per_chunk=B // 5 + 1, device=DEV)  # 5 chunks, the last one ragged
    for _ in range(2):  # buffers are reused across runs
        pipe.run(hb, hp, hr, ha, hc, temperature=0.0)
        torch.cuda.synchronize()
        pipe.raise_for_status()
        assert torch.equal(hc, res.counts.cpu()) and torch.equal(ha, act.cpu())
    pipe.close()
    # position-major form: host positions in, 21 (action, count) pairs per position out
    hb1, hp1 = torch.from_numpy(boards).pin_memory(), torch.from_numpy(players).pin_memory()
    ha.fill_(-7)
    hc.fill_(-7)
    pipe = bg.HostPipeline(w, items_per_chunk=len(boards) // 4 + 1, device=DEV, all_rolls=True)
    pipe.run(hb1, hp1, None, ha, hc, temperature=0.0)
    torch.cuda.synchronize()
    pipe.raise_for_status()
    assert torch.equal(hc, res.counts.cpu()) and torch.equal(ha, act.cpu())
    # every chunk's status is kept: an invalid board in the FIRST chunk is still reported after the last one
    bad = boards.copy()
    bad[3, 7] = 19
    pipe.run(torch.from_numpy(bad).pin_memory(), hp1, None, ha, hc, temperature=0.0)
    with pytest.raises(bg.BgError):
        pipe.raise_for_status()
    pipe.raise_for_status()  # the status word is reset by the read
    pipe.close()


def test_fused_movegen_eval_equals_separate_calls(bg, oracle):
    """bg_movegen_eval (evaluation of the bulk tier overlapping the tail tiers) == bg_movegen followed by bg_eval over the pool."""
    boards, players = oracle.random_positions(20000, seed=41)
    ib, ip, ir = oracle.all_rolls_items(boards, players)
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "values.npz"))
    w = bg.prepare_weights(torch.from_numpy(g["packed"]).to(DEV), int(g["H"]))
    db, dp, dr = dev(ib), dev(ip), dev(ir)
    ref = bg.movegen(db, dp, dr, item_cap=500, pool_cap=len(ib) * 30)
    v_ref = bg.evaluate(ref.boards, ref.flags, w, n_dev=ref.total_dev)
    cap = len(ib) * 30
    pool = torch.empty((cap, 52), dtype=torch.int8, device=DEV)
    flags = torch.empty(cap, dtype=torch.uint8, device=DEV)
    vals = torch.full((cap,), float("nan"), device=DEV)
    for _ in range(2):
        res, v = bg.movegen_evaluate(db, dp, dr, w, pool, flags, vals, check_status=True)
        torch.cuda.synchronize()
        assert res.total == ref.total and torch.equal(res.counts, ref.counts)
        # pool placement is run dependent: compare per item, in action order
        o1, b1, _ = ref.canonical()
        o2, b2, _ = res.canonical()
        assert torch.equal(o1, o2) and torch.equal(b1, b2)
        kept = torch.clamp(ref.counts.long(), max=500)
        item = torch.repeat_interleave(torch.arange(len(ib), device=DEV), kept)
        k = torch.arange(int(o1[-1]), device=DEV) - o1[item]
        assert torch.equal(v_ref[ref.offsets[item] + k], v[res.offsets[item] + k])  # same kernels, same rows: bit-identical
        assert not torch.isnan(v[: res.total]).any()
