"""Python prototype of the position-major move generator's ALGORITHM (csrc/movegen21.cu), checked against the C oracle.

Validates, on CPU and before any GPU time is spent, the three claims the CUDA kernel relies on:
  1. code equality <=> board equality (non-doubles: sorted sources / destinations after cancellation + intermediate hit;
     doubles: sorted multiset of source points);
  2. pruning candidates that provably duplicate an EARLIER candidate (commuting second-order pairs of a non-double; for doubles
     a slot below the parent's last slot whose move was already legal from the grandparent) never changes the
     first-occurrence order;
  3. breadth-first by level with first-occurrence dedup == the reference's DFS order.
Test infrastructure only (imports the oracle).
    python tests/tools/proto_movegen21.py [n_positions] [seed]
"""
import sys, os
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pyoracle as po

NONE = 31


class Root:
    def __init__(self, board, player):
        self.player = player
        own = board[0:24] if player == 0 else board[24:48]
        opp = board[24:48] if player == 0 else board[0:24]
        self.cnt = [int(x) for x in own]
        self.bar = int(board[48 + player])
        self.off = int(board[50 + player])
        self.blocked = sum(1 << i for i in range(24) if opp[i] >= 2)
        self.blot = sum(1 << i for i in range(24) if opp[i] == 1)
        self.home = 0xFC0000 if player == 0 else 0x3F
        self.valid15 = sum(self.cnt) + self.bar + self.off == 15
        self.dir = 1 if player == 0 else -1
        self.board = board


class Node:
    __slots__ = ("cnt", "bar", "off", "hit")

    def __init__(self, cnt, bar, off, hit):
        self.cnt, self.bar, self.off, self.hit = cnt, bar, off, hit


def move_mask(n, r, die):
    """slot mask (bits 0..23 point moves, 24 bar entry, 25 farthest bear-off, 26 exact bear-off), last"""
    occ = sum(1 << i for i in range(24) if n.cnt[i] > 0)
    if n.off == 15:
        return 0, 0
    if n.bar > 0:
        e = die - 1 if r.player == 0 else 24 - die
        return (0 if (r.blocked >> e) & 1 else 1 << 24), 0
    if r.player == 0:
        vm = occ & ~(r.blocked >> die) & ((1 << (24 - die)) - 1)
    else:
        vm = occ & ~(r.blocked << die) & (0xFFFFFF & ~((1 << die) - 1))
    last = 0
    if r.valid15 and (occ & ~r.home) == 0:
        if r.player == 0:
            last = (occ & -occ).bit_length() - 1 if occ else 18
            far = last + die >= 24
            ps = 24 - die
        else:
            last = occ.bit_length() - 1 if occ else 5
            far = last - die < 0
            ps = die - 1
        if far:
            vm |= 1 << 25
        if ps != last and (occ >> ps) & 1:
            vm |= 1 << 26
    return vm, last


def slot_se(r, slot, last, die):
    if slot < 24:
        return slot, slot + r.dir * die
    if slot == 24:
        return 24, (die - 1 if r.player == 0 else 24 - die)
    if slot == 25:
        return last, 25
    return (24 - die if r.player == 0 else die - 1), 25


def apply(n, r, s, e):
    cnt = list(n.cnt)
    bar, off, hit = n.bar, n.off, n.hit
    if s == 24:
        bar -= 1
    else:
        cnt[s] -= 1
    if e == 25:
        off += 1
    else:
        cnt[e] += 1
        if (r.blot & ~hit) >> e & 1:
            hit |= 1 << e
    return Node(cnt, bar, off, hit)


def node_board(n, r):
    b = r.board.copy()
    p = r.player
    o = 0 if p == 0 else 24
    q = 24 if p == 0 else 0
    b[o:o + 24] = n.cnt
    for i in range(24):
        if (n.hit >> i) & 1:
            b[q + i] -= 1
    b[48 + p] = n.bar
    b[48 + 1 - p] += bin(n.hit).count("1")
    b[50 + p] = n.off
    return b


def bits(m):
    i = 0
    while m:
        if m & 1:
            yield i
        m >>= 1
        i += 1


def nd_code(r, s1, e1, s2, e2):
    """non-doubles result code: sorted sources, sorted destinations after cancelling a point that is both, + intermediate hit"""
    inter = NONE
    if s2 == e1:  # a checker lands on e1 and a checker leaves e1: -[s1] +[e2]
        if (r.blot >> e1) & 1:
            inter = e1
        srcs, dsts = (s1, NONE), (e2, NONE)
    elif s1 == e2:  # the second move lands on the point the first one left
        srcs, dsts = (s2, NONE), (e1, NONE)
    else:
        srcs, dsts = tuple(sorted((s1, s2))), tuple(sorted((e1, e2)))
    return srcs + dsts + (inter,)


def code_board_nd(r, code):
    sa, sb, da, db, inter = code
    n = Node(list(r.cnt), r.bar, r.off, 0)
    for s in (sa, sb):
        if s == NONE:
            continue
        if s == 24:
            n.bar -= 1
        else:
            n.cnt[s] -= 1
    for d in (da, db):
        if d == NONE:
            continue
        if d == 25:
            n.off += 1
        else:
            n.cnt[d] += 1
            if (r.blot >> d) & 1:
                n.hit |= 1 << d
    if inter != NONE:
        n.hit |= 1 << inter
    return node_board(n, r)


def code_board_dbl(r, srcs, die):
    n = Node(list(r.cnt), r.bar, r.off, 0)
    for s in srcs:
        if s == NONE:
            continue
        if s == 24:
            n.bar -= 1
            e = die - 1 if r.player == 0 else 24 - die
        else:
            n.cnt[s] -= 1
            e = s + r.dir * die
            if e < 0 or e > 23:
                e = 25
        if e == 25:
            n.off += 1
        else:
            n.cnt[e] += 1
            if (r.blot >> e) & 1:
                n.hit |= 1 << e
    return node_board(n, r)


STATS = {"nd_cand": 0, "nd_cand_unpruned": 0, "nd_uniq": 0, "dbl_cand": 0, "dbl_cand_unpruned": 0, "dbl_uniq": 0, "dbl_dups_after_prune": 0,
         "nd_dups_after_prune": 0, "n1": [], "nd_cand_pos": [], "nd_parents_pos": [], "dbl_lvl": []}


def gen_position(board, player, prune=True):
    r = Root(board, player)
    root = Node(list(r.cnt), r.bar, r.off, 0)
    # level 1 for the six dice
    M1, L1, kids = {}, {}, {}
    for d in range(1, 7):
        m, last = move_mask(root, r, d)
        M1[d], L1[d] = m, last
        ks = []
        for slot in bits(m):
            s, e = slot_se(r, slot, last, d)
            ks.append((slot, s, e, apply(root, r, s, e)))
        kids[d] = ks
    STATS["n1"].append(sum(len(kids[d]) for d in kids))
    out = {}
    nd_c = nd_p = 0
    # ---- non-doubles ----
    for (a, b) in po.DICE_ROLLS:
        if a == b:
            continue
        hi, lo = max(a, b), min(a, b)
        m2 = {}
        for (fa, fb) in ((hi, lo), (lo, hi)):
            m2[fa] = [move_mask(k[3], r, fb) for k in kids[fa]]
        has2 = {fa: any(m for m, _ in m2[fa]) for fa in (hi, lo)}
        segs = []  # list of (kind, first die, second die)
        n_hi = len(kids[hi])
        if has2[hi]:
            segs.append(("two", hi, lo, 0))
            if has2[lo]:
                segs.append(("two", lo, hi, 1))
        else:
            if n_hi == 1:
                segs.append(("single", hi, lo, 0))
            elif n_hi == 0:
                segs.append(("two", lo, hi, 1) if has2[lo] else ("single", lo, hi, 1))
            else:
                if has2[lo]:
                    segs.append(("two", lo, hi, 1))
                else:
                    segs.append(("single", hi, lo, 0))
                    segs.append(("single", lo, hi, 1))
        seen, res = set(), []
        for kind, fa, fb, order in segs:
            for ci, (slot1, s1, e1, ch) in enumerate(kids[fa]):
                if kind == "single":
                    code = (s1, NONE, e1, NONE, NONE)
                    assert code not in seen
                    seen.add(code)
                    res.append(code)
                    continue
                m, last = m2[fa][ci]
                STATS["nd_cand_unpruned"] += bin(m).count("1")
                nd_p += 1
                if prune and order == 1 and r.bar == 0 and slot1 < 24 and has2[hi]:
                    # commuting pair: (s2 by hi, then s1 by lo) is an order-0 candidate with the same board
                    m &= ~(M1[hi] & 0xFFFFFF)
                for slot2 in bits(m):
                    s2, e2 = slot_se(r, slot2, last, fb)
                    code = nd_code(r, s1, e1, s2, e2)
                    STATS["nd_cand"] += 1
                    nd_c += 1
                    if code in seen:
                        STATS["nd_dups_after_prune"] += 1
                        continue
                    seen.add(code)
                    res.append(code)
        STATS["nd_uniq"] += len(res)
        out[(a, b)] = [code_board_nd(r, c) for c in res]
    STATS["nd_cand_pos"].append(nd_c)
    STATS["nd_parents_pos"].append(nd_p)
    # ---- doubles: BFS by level, node = (sorted sources, last slot t, Node) ----
    lv = []
    for d in range(1, 7):
        front = []
        for (slot, s, e, ch) in kids[d]:
            front.append(((s,), slot, ch))
        depth = 1 if front else 0
        sizes = [len(front)]
        while depth and depth < 4:
            seen, nxt = set(), []
            for (srcs, t, nd) in front:
                m, last = move_mask(nd, r, d)
                STATS["dbl_cand_unpruned"] += bin(m).count("1")
                if prune and t < 24:
                    keep = ~((1 << t) - 1)  # slots >= t
                    lt = t + r.dir * d  # landing point of the parent's last move
                    if 0 <= lt < t and nd.cnt[lt] == 1:
                        keep |= 1 << lt
                    m &= keep | (7 << 24)
                for slot in bits(m):
                    s, e = slot_se(r, slot, last, d)
                    code = tuple(sorted(srcs + (s,)))
                    STATS["dbl_cand"] += 1
                    if code in seen:
                        STATS["dbl_dups_after_prune"] += 1
                        continue
                    seen.add(code)
                    nxt.append((code, slot, apply(nd, r, s, e)))
            if not nxt:
                break
            front = nxt
            sizes.append(len(front))
            depth += 1
        lv.append(sizes)
        STATS["dbl_uniq"] += len(front) if depth else 0
        out[(d, d)] = [code_board_dbl(r, srcs, d) for (srcs, t, nd) in front] if depth else []
        # the rebuilt board must equal the incrementally applied node
        for (srcs, t, nd) in (front if depth else []):
            assert np.array_equal(code_board_dbl(r, srcs, d), node_board(nd, r))
    STATS["dbl_lvl"].append(lv)
    return out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 11
    po.build()
    b, p = po.random_positions(n, seed=seed)
    ib, ip, ir = po.all_rolls_items(b, p)
    off, ob, _ = po.movegen_batch(ib, ip, ir, want_moves=False)
    bad = 0
    for i in range(n):
        out = gen_position(b[i], int(p[i]))
        for j, roll in enumerate(po.DICE_ROLLS):
            want = ob[off[i * 21 + j]:off[i * 21 + j + 1]]
            got = np.array(out[roll], np.int8).reshape(-1, 52)
            if got.shape != want.shape or not np.array_equal(got, want):
                bad += 1
                if bad < 5:
                    print("MISMATCH pos", i, "player", p[i], "roll", roll, "want", want.shape[0], "got", got.shape[0])
                    print(b[i].tolist())
    S = STATS
    print(f"positions {n}: mismatching items {bad}")
    print(f"non-doubles: candidates {S['nd_cand']} (unpruned {S['nd_cand_unpruned']}), unique {S['nd_uniq']}, duplicates left {S['nd_dups_after_prune']}")
    print(f"doubles:     candidates {S['dbl_cand']} (unpruned {S['dbl_cand_unpruned']}), unique final {S['dbl_uniq']}, duplicates left {S['dbl_dups_after_prune']}")
    for name in ("n1", "nd_cand_pos", "nd_parents_pos"):
        x = np.array(S[name])
        print(name, "mean %.1f p50 %d p90 %d p99 %d max %d" % (x.mean(), *np.percentile(x, [50, 90, 99]), x.max()))
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
