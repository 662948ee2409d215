"""GPU check of the position-major move generator (bg_movegen_all_rolls) against the C oracle and against the per-item kernels,
plus timing.  Test infrastructure (imports the oracle).
    python tests/tools/check_movegen21.py [n_oracle_positions] [n_timing_positions]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mlp_ppo_2ply_multi_b200 as bg
from oracle import pyoracle as po


def checksum(res, n_items):
    """order-sensitive 64-bit checksum per item of its afterstate list (device)"""
    kept = torch.clamp(res.counts.to(torch.int64), max=res.item_cap)
    off = torch.zeros(n_items + 1, dtype=torch.int64, device=kept.device)
    off[1:] = torch.cumsum(kept, 0)
    T = int(off[-1].item())
    item = torch.repeat_interleave(torch.arange(n_items, device=kept.device), kept, output_size=T)
    rank = torch.arange(T, device=kept.device) - off[item]
    rows = res.offsets[item] + rank
    w = torch.arange(1, 53, device=kept.device, dtype=torch.int64) * 1000003
    out = torch.zeros(n_items, dtype=torch.int64, device=kept.device)
    for c0 in range(0, T, 1 << 24):  # chunked: [T,52] int64 does not fit for the full configuration
        sl = slice(c0, min(T, c0 + (1 << 24)))
        b = res.boards[rows[sl]].to(torch.int64)
        h = ((b + 17) * w).sum(1) * (rank[sl] * 2 + 1) * 0x9E3779B1
        out.index_add_(0, item[sl], h)
    return out, kept


def main():
    n_or = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    n_big = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
    dev = torch.device("cuda:0")
    po.build()
    b, p = po.random_positions(n_or, seed=4242)
    ib, ip, ir = po.all_rolls_items(b, p)
    off, ob, _ = po.movegen_batch(ib, ip, ir, want_moves=False)
    tb, tp = torch.from_numpy(b).to(dev), torch.from_numpy(p).to(dev)
    res = bg.movegen_all_rolls(tb, tp, item_cap=4096, pool_cap=int(off[-1]) + 1024)
    coff, cb, _ = res.canonical()
    ok_off = np.array_equal(coff.cpu().numpy(), off)
    ok_b = ok_off and np.array_equal(cb.cpu().numpy(), ob)
    print(f"oracle check: {n_or} positions x 21 rolls, {int(off[-1])} afterstates: offsets {'OK' if ok_off else 'DIFFER'}, boards {'OK' if ok_b else 'DIFFER'}", flush=True)
    if not ok_b:
        g = coff.cpu().numpy()
        cnt_g, cnt_o = np.diff(g), np.diff(off)
        bad = np.nonzero(cnt_g != cnt_o)[0]
        print("items with a different count:", len(bad), bad[:10], cnt_g[bad[:10]], cnt_o[bad[:10]])
        if ok_off:
            gb = cb.cpu().numpy()
            rows = np.nonzero((gb != ob).any(1))[0]
            it = np.searchsorted(off, rows, side="right") - 1
            print("rows differing:", len(rows), "items:", np.unique(it)[:10])
            i = int(it[0])
            print("item", i, "pos", i // 21, "roll", po.DICE_ROLLS[i % 21], "player", p[i // 21], b[i // 21].tolist())
            print("want", ob[off[i]:off[i + 1]].tolist()[:3])
            print("got ", gb[off[i]:off[i + 1]].tolist()[:3])
        else:
            i = int(bad[0])
            print("item", i, "pos", i // 21, "roll", po.DICE_ROLLS[i % 21], "player", p[i // 21], b[i // 21].tolist())
        return 1
    # flags / owner
    res2 = bg.movegen_all_rolls(tb, tp, item_cap=4096, pool_cap=int(off[-1]) + 1024, want_owner=True)
    T = int(res2.total)
    own = res2.owner[:T].to(torch.int64)
    assert bool((res2.flags[:T] == tp[own // 21]).all())
    assert bool(((res2.offsets[own] <= torch.arange(T, device=dev)) & (torch.arange(T, device=dev) < res2.offsets[own] + res2.counts[own])).all())
    print("flags / owner OK", flush=True)
    # larger: against the per-item kernels through an order-sensitive checksum, and timing
    ar_seed = 99
    from bench import make_positions
    boards, players = make_positions(bg, n_big, dev, ar_seed)
    P = boards.shape[0]
    rolls = torch.tensor(bg.DICE_ROLLS, dtype=torch.uint8, device=dev)
    ib = boards.repeat_interleave(21, dim=0).contiguous()
    ipl = players.repeat_interleave(21).contiguous()
    irl = rolls.repeat(P, 1).contiguous()
    pool_cap = P * 21 * 26 + (1 << 20)
    pool = torch.empty((pool_cap, 52), dtype=torch.int8, device=dev)
    flags = torch.empty(pool_cap, dtype=torch.uint8, device=dev)
    ws = torch.empty(bg._lib.lib().bg_movegen_workspace_bytes(21 * P), dtype=torch.uint8, device=dev)
    r_old = bg.movegen(ib, ipl, irl, item_cap=500, out_boards=pool, workspace=ws, want_owner=False, out_flags=flags)
    c_old, k_old = checksum(r_old, 21 * P)
    tot_old = r_old.total
    pool2 = torch.empty((pool_cap, 52), dtype=torch.int8, device=dev)
    r_new = bg.movegen_all_rolls(boards, players, item_cap=500, out_boards=pool2, workspace=ws, out_flags=flags)
    c_new, k_new = checksum(r_new, 21 * P)
    same_cnt = bool((r_old.counts == r_new.counts).all())
    same_sum = bool((c_old == c_new).all())
    print(f"per-item kernels vs position-major, {P} positions ({21 * P} items, {tot_old} afterstates): counts {'OK' if same_cnt else 'DIFFER'}, "
          f"order-sensitive checksums {'OK' if same_sum else 'DIFFER'}, totals {tot_old} / {r_new.total}", flush=True)
    ovf = torch.frombuffer(ws[32:48].cpu().numpy().tobytes(), dtype=torch.int32) if False else ws[32:48].view(torch.int32).cpu()
    print("overflow-list counts after the position-major call (lists 1-4):", ovf.tolist(), flush=True)
    for name, fn in (("per-item", lambda: bg.movegen(ib, ipl, irl, item_cap=500, out_boards=pool, workspace=ws, want_owner=False, out_flags=flags, check_status=False)),
                     ("position-major", lambda: bg.movegen_all_rolls(boards, players, item_cap=500, out_boards=pool2, workspace=ws, out_flags=flags, check_status=False))):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        R = 5
        for _ in range(R):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / R
        print(f"{name}: {ms:.2f} ms per pass, {tot_old / ms / 1e6:.2f} G afterstates/s, {21 * P / ms / 1e3:.1f} M items/s", flush=True)
    return 0 if (same_cnt and same_sum) else 1


if __name__ == "__main__":
    sys.exit(main())
