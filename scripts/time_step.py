"""Timing driver: the config-2 step in its board and compact forms (CUDA events, per pass)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mlp_ppo_2ply_multi_b200 as bg
from bench import H, make_positions, packed_random_weights

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
dev = torch.device("cuda:0")
if os.environ.get("BG_TILES"):  # development: force the evaluator's tile schedule (0 static, 1 dynamic)
    bg._lib.lib().bg_eval_tc_tile_schedule(int(os.environ["BG_TILES"]))
boards, players = make_positions(bg, n, dev, 2026)
P = boards.shape[0]
pool_cap = P * 21 * 26 + (1 << 20)
pool = torch.empty((pool_cap, 52), dtype=torch.int8, device=dev)
codes = torch.empty(pool_cap, dtype=torch.int64, device=dev)
flags = torch.empty(pool_cap, dtype=torch.uint8, device=dev)
values = torch.empty(pool_cap, dtype=torch.float32, device=dev)
ws = torch.empty(bg._lib.lib().bg_movegen_workspace_bytes(21 * P), dtype=torch.uint8, device=dev)
w = bg.prepare_weights(packed_random_weights(0).to(dev), H)


def timeit(fn, reps=5):
    for _ in range(2):
        r = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r


t, r = timeit(lambda: bg.movegen_all_rolls(boards, players, item_cap=500, out_boards=pool, workspace=ws, out_flags=flags, check_status=False))
tot = r.total
print(f"movegen (boards):  {t:.2f} ms   {tot / t / 1e6:.2f} G afterstates/s")
t, r = timeit(lambda: bg.movegen_all_rolls_compact(boards, players, None, item_cap=500, out_codes=codes, workspace=ws))
print(f"movegen (compact): {t:.2f} ms   {r.total / t / 1e6:.2f} G afterstates/s")
t, _ = timeit(lambda: bg.evaluate(pool, flags, w, n_dev=r.total_dev, out=values))
print(f"eval (boards):     {t:.2f} ms")
t, _ = timeit(lambda: bg.evaluate_codes(r, w, out=values))
print(f"eval (compact):    {t:.2f} ms")
t, _ = timeit(lambda: bg.movegen_evaluate_all_rolls(boards, players, w, pool, flags, values, workspace=ws, item_cap=500))
print(f"step (boards):     {t:.2f} ms   {tot / t / 1e6:.2f} G afterstates/s")
t, _ = timeit(lambda: bg.movegen_all_rolls_compact(boards, players, w, item_cap=500, out_codes=codes, out_values=values, workspace=ws))
print(f"step (compact):    {t:.2f} ms   {tot / t / 1e6:.2f} G afterstates/s")
