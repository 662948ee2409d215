"""Developer micro-benchmark (NOT bench.py): times bg_movegen / bg_eval on oracle-generated positions."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import mlp_ppo_2ply_multi_b200 as bg
from oracle import pyoracle as po

n_pos = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
dev = torch.device("cuda:0")
t = time.time(); boards, players = po.random_positions(n_pos, seed=2026); print("positions", time.time() - t, "s", os.cpu_count(), "cpus")
ib, ip, ir = po.all_rolls_items(boards, players)
B = len(ib)
dib, dip, dir_ = (torch.from_numpy(x).to(dev) for x in (ib, ip, ir))
g = np.load("tests/golden/values.npz"); w = bg.prepare_weights(torch.from_numpy(g["packed"]).to(dev), 128)
pool_cap = B * 26
out_boards = torch.empty((pool_cap, 52), dtype=torch.int8, device=dev)
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for rolls_name, sel in (("all21", slice(None)), ("nondoubles", None), ("doubles", None)):
    if sel is None:
        m = (ir[:, 0] != ir[:, 1]) if rolls_name == "nondoubles" else (ir[:, 0] == ir[:, 1])
        idx = torch.from_numpy(np.nonzero(m)[0]).to(dev)
        b_, p_, r_ = dib[idx].contiguous(), dip[idx].contiguous(), dir_[idx].contiguous()
    else:
        b_, p_, r_ = dib, dip, dir_
    for it in range(3):
        e0 = ev(); res = bg.movegen(b_, p_, r_, item_cap=4096, out_boards=out_boards, check_status=False); e1 = ev()
        v = bg.evaluate(out_boards, res.flags, w, n_dev=res.total_dev); e2 = ev()
        torch.cuda.synchronize()
    tot = res.total
    tm, te = e0.elapsed_time(e1), e1.elapsed_time(e2)
    print(f"{rolls_name}: items {len(b_)} afterstates {tot} status {int(res.status_dev.item())} movegen {tm:.2f} ms ({len(b_)/tm/1e3:.2f} M items/s, {tot/tm/1e3:.1f} M after/s) "
          f"eval {te:.2f} ms ({tot/te/1e3:.1f} M/s) both {tot/(tm+te)/1e3:.1f} M after/s")
ws = np.frombuffer(bg.ops._workspaces[0][:32].cpu().numpy().tobytes(), np.int32)
print("tier2 items", ws[6], "tier3 items", ws[7])
