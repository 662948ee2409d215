"""dev: per-phase clock64 stamps of k_eval_tc on the bench workload (compact pool, 65,536 positions x 21 rolls).

Needs a library built with the stamps compiled in (they are off in the shipped build):
    BG_OUT=/root/repo/build/libbgarena_stamps.so BG_NVCC_EXTRA="-DBG_TC_STAMPS=1 -DBG_TC_STAMP_FIRST=6000" bash mlp-ppo-2ply-multi_b200/csrc/build.sh
    BG_LIBBGARENA=build/libbgarena_stamps.so python scripts/dev_tc_stamps.py
Prints, for CTA 0 and 16 consecutive local tiles, the cycle offsets of each phase of builder warp 0, of the epilogue warp (lane
quarter 0) that owns the tile and of the MMA issuer (DESIGN.md section 4.4 quotes these)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mlp_ppo_2ply_multi_b200 as bg  # noqa: E402
from bench import H, make_positions, packed_random_weights  # noqa: E402

dev = torch.device("cuda:0")
boards, players = make_positions(bg, 1048576, dev, 2026)
w = bg.prepare_weights(packed_random_weights(0).to(dev), H)
r = bg.movegen_all_rolls_compact(boards, players, None, item_cap=500)
out = torch.empty(r.codes.numel(), dtype=torch.float32, device=dev)
for _ in range(3):  # the evaluator alone, one launch over the whole compact pool
    bg.evaluate_codes(r, w, out=out)
torch.cuda.synchronize()
buf = np.zeros((3, 64, 8), np.int64)
L = bg._lib.lib()
L.bg_dev_tc_stamps.argtypes = [C.c_void_p]
print("rc", L.bg_dev_tc_stamps(buf.ctypes.data), "afterstates", int(r.total))
bld, epi, iss = buf[0], buf[1], buf[2]
print("builder warp 0, per local tile: [start, A[slot] free, row written + arrived, tile n+1 rebuilt, loads of n+2 issued, index load of n+3 issued], period")
for t in range(8, 24):
    s = bld[t]
    print(t, [int(x - s[0]) for x in s[:6]], "period (two tiles)", int(bld[t + 2][0] - s[0]))
print("epilogue warp of lane quarter 0 that owns the tile: [start, mma_done seen, D in registers (d_free), value stored]; period = two tiles")
for t in range(8, 24):
    s = epi[t]
    print(t, [int(x - s[0]) for x in s[:4]], "period", int(epi[t + 2][0] - s[0]))
print("issuer: [a_full seen, d_free seen, 26 MMAs + commit issued], period; and the offset of this tile's a_full from the builder's arrive")
for t in range(8, 24):
    s = iss[t]
    print(t, [int(x - s[0]) for x in s[:3]], "period", int(iss[t + 1][0] - s[0]), "after arrive", int(s[0] - bld[t][2]))
per = np.diff(iss[:, 0])
print("mean period (cycles per tile):", float(per[4:60].mean()))
