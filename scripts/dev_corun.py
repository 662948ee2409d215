"""dev: do the position-major move generator and the tcgen05 evaluator overlap when they run on two streams?"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mlp_ppo_2ply_multi_b200 as bg
from bench import H, make_positions, packed_random_weights

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
boards, players = make_positions(bg, n, dev, 2026)
P = boards.shape[0]
pool_cap = P * 21 * 26 + (1 << 20)
poolA = torch.empty((pool_cap, 52), dtype=torch.int8, device=dev)
poolB = torch.empty((pool_cap, 52), dtype=torch.int8, device=dev)
flagsA = torch.empty(pool_cap, dtype=torch.uint8, device=dev)
flagsB = torch.empty(pool_cap, dtype=torch.uint8, device=dev)
values = torch.empty(pool_cap, dtype=torch.float32, device=dev)
ws = torch.empty(bg._lib.lib().bg_movegen_workspace_bytes(21 * P), dtype=torch.uint8, device=dev)
w = bg.prepare_weights(packed_random_weights(0).to(dev), H)
rA = bg.movegen_all_rolls(boards, players, item_cap=500, out_boards=poolA, workspace=ws, out_flags=flagsA)
nA = rA.total
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream(priority=-1)


def timeit(fn, reps=4):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def gen():
    bg.movegen_all_rolls(boards, players, item_cap=500, out_boards=poolB, workspace=ws, out_flags=flagsB, check_status=False)


def ev():
    bg.evaluate(poolA[:nA], flagsA[:nA], w, out=values)


def both():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur)
    s2.wait_stream(cur)
    with torch.cuda.stream(s2):
        ev()
    with torch.cuda.stream(s1):
        gen()
    cur.wait_stream(s1)
    cur.wait_stream(s2)


for ctas in sys.argv[2:] or ["5", "3", "2", "1"]:
    os.environ["BG_MG21_CTAS"] = ctas
    tg, te, tb = timeit(gen), timeit(ev), timeit(both)
    print(f"movegen CTAs/SM {ctas}: movegen {tg:.2f} ms, eval {te:.2f} ms, sum {tg + te:.2f}, both on two streams {tb:.2f} ms", flush=True)
