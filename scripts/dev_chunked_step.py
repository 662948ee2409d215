"""dev: the config-2 step run as C back-to-back calls over 1/C of the positions each (does alternating the ~0.7 kW move generator with the
~1.2 kW evaluator at a finer grain keep the SM clock up under the 1 kW cap?)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mlp_ppo_2ply_multi_b200 as bg  # noqa: E402
from bench import H, ClockSampler, make_positions, packed_random_weights  # noqa: E402

dev = torch.device("cuda:0")
n = 1048576
boards, players = make_positions(bg, n, dev, 2026)
w = bg.prepare_weights(packed_random_weights(0).to(dev), H)
for C in (1, 2, 4, 8, 16, 1):
    P = n // C
    cap = P * 21 * 26 + (1 << 20)
    codes = [torch.empty(cap, dtype=torch.int64, device=dev) for _ in range(min(C, 2))]
    vals = [torch.empty(cap, dtype=torch.float32, device=dev) for _ in range(min(C, 2))]
    ws = torch.empty(bg._lib.lib().bg_movegen_workspace_bytes(21 * P), dtype=torch.uint8, device=dev)

    def step():
        tot = 0
        for c in range(C):
            r, _ = bg.movegen_all_rolls_compact(boards[c * P:(c + 1) * P], players[c * P:(c + 1) * P], w, item_cap=500, out_codes=codes[c & 1],
                                                out_values=vals[c & 1], workspace=ws)
        return r

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    smp = ClockSampler(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"C={C:2d}: {ms:.2f} ms per step  {smp.stop()}", flush=True)
    del codes, vals, ws
    torch.cuda.empty_cache()
