"""Developer check of the tcgen05 evaluator: values vs the double-accumulated oracle and vs the FFMA kernel, timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import mlp_ppo_2ply_multi_b200 as bg
from oracle import pyoracle as po

dev = torch.device("cuda:0")
g = np.load("tests/golden/values.npz")
boards, players = po.random_positions(3000, seed=12)
ib, ip, ir = po.all_rolls_items(boards, players)
off, ob, _ = po.movegen_batch(ib, ip, ir, want_moves=False)
flags = np.repeat(ip, np.diff(off))
print("rows", len(ob))
for which in ("packed", "packed_init0"):
    w = bg.prepare_weights(torch.from_numpy(g[which]).to(dev), 128)
    d_b, d_f = torch.from_numpy(ob).to(dev), torch.from_numpy(flags).to(dev)
    v = bg.evaluate(d_b, d_f, w)
    torch.cuda.synchronize()
    st = bg._lib.lib().bg_eval_tc_status()
    ref = po.value(g[which], 128, ob[:200000], flags[:200000])
    err = np.abs(v.cpu().numpy()[:200000] - ref)
    print(which, "tc status", st, "max err", err.max(), "mean err", err.mean(), "n>1e-5", int((err > 1e-5).sum()))
    # small batch goes through the FFMA kernel
    v2 = torch.cat([bg.evaluate(d_b[i:i + 30000], d_f[i:i + 30000], w) for i in range(0, 200000, 30000)])
    print("   ffma vs oracle", np.abs(v2.cpu().numpy()[:200000] - ref[:len(v2)]).max(), " tc vs ffma", (v[:len(v2)] - v2).abs().max().item())
big = torch.from_numpy(np.tile(ob, (30, 1))[:20_000_000]).to(dev)
bf = torch.from_numpy(np.tile(flags, 30)[:20_000_000]).to(dev)
out = torch.empty(len(big), device=dev)
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); bg.evaluate(big, bf, w, out=out); e1.record(); torch.cuda.synchronize()
print("tc eval", len(big) / e0.elapsed_time(e1) / 1e6, "G boards/s;", e0.elapsed_time(e1), "ms")
