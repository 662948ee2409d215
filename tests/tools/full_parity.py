#!/usr/bin/env python
"""BASELINE configs[1] at FULL size: 1,048,576 random-legal positions (oracle playouts, seed 2026) x the 21 rolls = 22,020,096 items.
Every item's legal-afterstate list from bg_movegen is compared with the CPU oracle -- count, content and ORDER -- through a 64-bit
order-sensitive checksum per item; features are checked bit-exactly through bg_encode on a strided sample of the afterstates and
values through bg_eval to 1e-5.  Runs on the GPU box in chunks of 131,072 positions (the oracle side is the slow part).
    python tests/tools/full_parity.py [n_positions [positions_per_call]] > profiles/r01_full_parity.txt"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mlp_ppo_2ply_multi_b200 as bg  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

DEV = "cuda:0"
COEF = (np.arange(1, 14, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) | np.uint64(1)


def oracle_checksums(o_off, o_boards, n_items):
    total = int(o_off[-1])
    words = np.ascontiguousarray(o_boards).view(np.uint32).reshape(-1, 13)
    counts = np.diff(o_off)
    k = np.arange(total, dtype=np.int64) - np.repeat(o_off[:-1], counts)
    contrib = np.empty(total, np.uint64)
    CH = 1 << 22
    for lo in range(0, total, CH):
        hi = min(total, lo + CH)
        h = (words[lo:hi].astype(np.uint64) * COEF).sum(1, dtype=np.uint64)
        h ^= h >> np.uint64(29)
        contrib[lo:hi] = h * (2 * k[lo:hi] + 1).astype(np.uint64)
    want = np.zeros(n_items, np.uint64)
    nz = counts > 0
    want[nz] = np.add.reduceat(contrib, o_off[:-1][nz])
    return want


def gpu_checksums(res, total, n_items):
    got = torch.zeros(n_items, dtype=torch.int64, device=DEV)
    coef_t = torch.from_numpy(COEF.view(np.int64)).to(DEV)
    gw = res.boards[:total].view(torch.int32).reshape(-1, 13)
    own = res.owner[:total].long()
    CH = 1 << 22
    for lo in range(0, total, CH):
        hi = min(total, lo + CH)
        w64 = gw[lo:hi].to(torch.int64) & 0xFFFFFFFF
        h = (w64 * coef_t).sum(1)
        h = h ^ ((h >> 29) & ((1 << 35) - 1))
        kk = torch.arange(lo, hi, device=DEV) - res.offsets[own[lo:hi]]
        got.index_add_(0, own[lo:hi], h * (2 * kk + 1))
    return got.cpu().numpy().view(np.uint64)


def main():
    n_pos = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
    g = np.load(os.path.join(ROOT, "tests", "golden", "values.npz"))
    packed, H = g["packed"], int(g["H"])
    w = bg.prepare_weights(torch.from_numpy(packed).to(DEV), H)
    t0 = time.time()
    boards, players = po.random_positions(n_pos, seed=2026)
    CHUNK = int(sys.argv[2]) if len(sys.argv) > 2 else 262144  # 262,144 positions x 21 = 5.5 M items per call: every capacity tier is exercised
    items = afterstates = bad_items = 0
    feat_rows = feat_bad = 0
    max_dv = 0.0
    for lo in range(0, n_pos, CHUNK):
        ib, ip, ir = po.all_rolls_items(boards[lo:lo + CHUNK], players[lo:lo + CHUNK])
        o_off, o_boards, _ = po.movegen_batch(ib, ip, ir, want_moves=False)
        total = int(o_off[-1])
        res = bg.movegen(*(torch.from_numpy(x).to(DEV) for x in (ib, ip, ir)), item_cap=4096, pool_cap=total + 4096, want_owner=True)
        ok = res.total == total and np.array_equal(res.counts.cpu().numpy(), np.diff(o_off))
        bad = int((gpu_checksums(res, total, len(ib)) != oracle_checksums(o_off, o_boards, len(ib))).sum()) if ok else len(ib)
        items += len(ib)
        afterstates += total
        bad_items += bad
        # features (bit-exact) and values (1e-5) on every 97th afterstate of the oracle's CSR list
        sel = np.arange(0, total, 97)
        sb = np.ascontiguousarray(o_boards[sel])
        sf = np.repeat(ip, np.diff(o_off))[sel].astype(np.uint8)
        f_gpu = bg.encode(torch.from_numpy(sb).to(DEV), torch.from_numpy(sf).to(DEV)).cpu().numpy()
        f_ref = po.encode(sb, sf)
        feat_rows += len(sel)
        feat_bad += int((f_gpu.view(np.uint32) != f_ref.view(np.uint32)).any(1).sum())
        v_gpu = bg.evaluate(torch.from_numpy(sb).to(DEV), torch.from_numpy(sf).to(DEV), w).cpu().numpy()
        max_dv = max(max_dv, float(np.abs(v_gpu - po.value(packed, H, sb, sf)).max()))
        print(f"positions {lo:8d}..{min(lo + CHUNK, n_pos):8d}: items {len(ib):8d} afterstates {total:10d} mismatching items {bad} "
              f"feature rows {len(sel)} max|dV| so far {max_dv:.2e}  [{time.time() - t0:.0f} s]", flush=True)
    print(f"TOTAL: {n_pos} positions x 21 rolls = {items} items, {afterstates} afterstates; items whose (count, content, order) checksum differs "
          f"from the oracle: {bad_items}; feature rows compared bit-exactly: {feat_rows}, differing: {feat_bad}; max |dV| = {max_dv:.3e} (tolerance 1e-5)")
    sys.exit(0 if bad_items == 0 and feat_bad == 0 and max_dv < 1e-5 else 1)


if __name__ == "__main__":
    main()
