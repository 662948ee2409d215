#!/usr/bin/env python
"""BASELINE configs[1] at FULL size, the VALUE side of the headline step: 1,048,576 positions x 21 rolls through
bg_movegen_eval_all_rolls_compact (compact pool, afterstates rebuilt inside the tcgen05 evaluator).  Checks, per chunk of positions:
  * every value of the compact step against the board-pool step (bg_movegen_eval_all_rolls: boards written to HBM and read back) -- the
    two pools list an item's afterstates in the same order, so the comparison is row by row through the items' offsets;
  * every 499th row: the afterstate materialised from its code (bg_afterstates_from_codes) evaluated by the double-accumulated CPU oracle
    (1e-5 contract) and by the CUDA-core FFMA kernel;
  * the evaluator's status word.
    python tests/tools/full_values.py [n_positions [positions_per_call]] > profiles/r02_full_values.txt"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mlp_ppo_2ply_multi_b200 as bg  # noqa: E402
from bench import make_positions  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

DEV = torch.device("cuda:0")


def main():
    n_pos = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
    CH = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
    g = np.load(os.path.join(ROOT, "tests", "golden", "values.npz"))
    packed, H = g["packed"], int(g["H"])
    w = bg.prepare_weights(torch.from_numpy(packed).to(DEV), H)
    boards, players = make_positions(bg, n_pos, DEV, 2026)  # the benchmark's positions
    t0 = time.time()
    rows_total = sample_total = 0
    d_pool = d_oracle = d_ffma = 0.0
    for lo in range(0, n_pos, CH):
        b, p = boards[lo:lo + CH], players[lo:lo + CH]
        P = b.shape[0]
        cap = P * 21 * 26 + (1 << 20)
        res, vals = bg.movegen_all_rolls_compact(b, p, w, item_cap=500, pool_cap=cap, check_status=True)
        T = int(res.total)
        pool = torch.empty((cap, 52), dtype=torch.int8, device=DEV)
        flags = torch.empty(cap, dtype=torch.uint8, device=DEV)
        vb = torch.empty(cap, dtype=torch.float32, device=DEV)
        rb, vb = bg.movegen_evaluate_all_rolls(b, p, w, pool, flags, vb, item_cap=500)
        assert int(rb.total) == T and torch.equal(rb.counts, res.counts)
        kept = torch.clamp(res.counts.to(torch.int64), max=500)
        start = torch.cumsum(kept, 0) - kept
        item = torch.repeat_interleave(torch.arange(kept.numel(), device=DEV), kept)
        k = torch.arange(T, device=DEV) - start[item]
        rc, rbp = res.offsets[item] + k, rb.offsets[item] + k
        d_pool = max(d_pool, (vals[rc] - vb[rbp]).abs().max().item())
        sel = rc[::499]
        ab = res.afterstates(sel)
        fl = p[(res.codes[sel] >> 32).to(torch.int64)]
        want = po.value(packed, H, ab.cpu().numpy(), fl.cpu().numpy())
        d_oracle = max(d_oracle, float(np.abs(vals[sel].cpu().numpy() - want).max()))
        v_ff = torch.cat([bg.evaluate(ab[i:i + 30000], fl[i:i + 30000], w) for i in range(0, ab.shape[0], 30000)])
        d_ffma = max(d_ffma, (vals[sel] - v_ff).abs().max().item())
        st = bg._lib.lib().bg_eval_tc_status()
        rows_total += T
        sample_total += int(sel.numel())
        print(f"positions {lo:8d}..{lo + P:8d}: rows {T:10d}  max|compact - board pool| {d_pool:.2e}  sample {int(sel.numel()):7d}: max|dV| vs oracle {d_oracle:.2e}, "
              f"vs FFMA kernel {d_ffma:.2e}  evaluator status {st}  [{time.time() - t0:.0f} s]", flush=True)
        assert st == 0
        del pool, flags, vb, rb, res, vals, item, k, rc, rbp
        torch.cuda.empty_cache()
    print(f"TOTAL: {n_pos} positions x 21 rolls, {rows_total} afterstate values of the compact step; max |compact - board-pool step| = {d_pool:.3e}; "
          f"{sample_total} sampled rows: max |dV| vs the oracle = {d_oracle:.3e} (tolerance 1e-5), vs the FFMA kernel = {d_ffma:.3e}")
    sys.exit(0 if d_oracle < 1e-5 and d_pool < 2e-6 and d_ffma < 2e-6 else 1)


if __name__ == "__main__":
    main()
