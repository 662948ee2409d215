#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native backgammon hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--positions P]

Workload (BASELINE.json configs[1]): P = 1,048,576 random-but-legal positions x the 21 unordered rolls
(22,020,096 (position, roll) items, ~4.8e8 afterstates) per GPU.  One "step" = one pass of the hot path over that batch:
legal-move generation (bg_movegen_all_rolls: one warp per position, all 21 rolls) + fused 198-feature encode + sigmoid-MLP value of every
afterstate (bg_eval), one C-ABI call (bg_movegen_eval_all_rolls).
Metric: afterstates evaluated per second (whole job, all GPUs).  Positions are synthetic: produced by the arena itself
playing uniformly random legal moves from the start position (temperature -> infinity), snapshotted at spread-out plies.
The first 65,536 positions are the committed fixture tests/golden/bench_positions.npz (made by this generator, scripts/make_bench_positions.py),
which is also what `--impl reference` and the cpu_baseline leg run on.
Also reported: `e2e` (bg_hostpipe_run: host positions in, greedy action + count per item out, copies inside the timed region),
`roofline` for the dominant kernel, `cpu_baseline` (the C oracle port on the host cores, bounded sample),
`selfplay_1ply` (BASELINE configs[2]: 65,536 concurrent games, games/s).
Under torchrun each rank owns an independent shard (weak scaling, no data-path collective); rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = 128
FLOP_PER_AFTERSTATE = 2 * 198 * H + 2 * H  # 50,944 dense FLOP per evaluated afterstate


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0]))
                mx.append(float(c[1]))
                power.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            busy = [s for s, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
            out = {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "power_w_max": max(power),
                   "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def packed_random_weights(seed=0):
    """Xavier-uniform 198->128->1 value net (reference policy_network.py:36-51), packed [W1^T | b1 | w2 | b2]."""
    import torch

    g = torch.Generator().manual_seed(seed)
    a1 = (6.0 / (198 + H)) ** 0.5
    W1 = (torch.rand(H, 198, generator=g) * 2 - 1) * a1
    b1 = (torch.rand(H, generator=g) * 2 - 1) / 198 ** 0.5
    a2 = (6.0 / (H + 1)) ** 0.5
    w2 = (torch.rand(1, H, generator=g) * 2 - 1) * a2
    b2 = (torch.rand(1, generator=g) * 2 - 1) / H ** 0.5
    return torch.cat([W1.t().contiguous().reshape(-1), b1, w2.reshape(-1), b2]).float()


def make_positions(bg, n_pos, device, seed):
    """n_pos reachable positions from the arena playing uniformly random legal moves (temperature 1e9)."""
    import torch

    G = min(131072, n_pos)
    snaps = (n_pos + G - 1) // G
    ar = bg.Arena(G, hidden_size=H, device=device, seed=seed, ring_experiences=1 << 22, ring_episodes=1 << 17)
    ar.set_weights(packed_random_weights(1).to(device), version=1, temperature=1e9)
    ar.reset()
    boards, players = [], []
    have = 0
    ar.step(97)
    for s in range(4 * snaps + 4):
        ar.step(13)
        while ar.drain(max_episodes=1 << 16, max_experiences=1 << 21).n_episodes:  # keep the ring empty; episodes are discarded
            pass
        b, p, _, st = ar.state()
        keep = st == 0
        boards.append(b[keep])
        players.append(p[keep])
        have += int(keep.sum().item())
        if have >= n_pos:
            break
    ar.close()
    boards = torch.cat(boards)[:n_pos].contiguous()
    players = torch.cat(players)[:n_pos].contiguous()
    return boards, players


FIXTURE = os.path.join(ROOT, "tests", "golden", "bench_positions.npz")


def fixture_positions():
    """the committed head of the benchmark's position set (numpy), or None"""
    if not os.path.exists(FIXTURE):
        return None
    import numpy as np

    g = np.load(FIXTURE)
    return g["boards"], g["players"]


def bench_positions(bg, n_pos, device, seed, rank):
    """n_pos positions: the fixture first (rank 0 only: the other ranks' shards are disjoint arena draws), then fresh arena draws"""
    import torch

    fx = fixture_positions() if rank == 0 else None
    if fx is None:
        b, p = make_positions(bg, n_pos, device, seed)
        return b, p, 0
    fb, fp = torch.from_numpy(fx[0]).to(device), torch.from_numpy(fx[1]).to(device)
    if n_pos <= fb.shape[0]:
        return fb[:n_pos].contiguous(), fp[:n_pos].contiguous(), n_pos
    b, p = make_positions(bg, n_pos - fb.shape[0], device, seed)
    return torch.cat([fb, b]).contiguous(), torch.cat([fp, p]).contiguous(), int(fb.shape[0])


def expand_rolls(bg, boards, players):
    import torch

    n = boards.shape[0]
    rolls = torch.tensor(bg.DICE_ROLLS, dtype=torch.uint8, device=boards.device)
    ib = boards.repeat_interleave(21, dim=0)
    ip = players.repeat_interleave(21)
    ir = rolls.repeat(n, 1)
    return ib.contiguous(), ip.contiguous(), ir.contiguous()


def run_python_reference(mode, cores):
    """oracle/reference_bench.py legs as subprocesses (the reference needs its own sys.path and `spawn`); each prints one JSON line"""
    if mode == "off":
        return {"unavailable": "--pyref off"}
    games, secs = (100, 60) if mode == "full" else (30, 20)
    script = os.path.join(ROOT, "oracle", "reference_bench.py")
    legs = {"single_xavier_T1.5": ["single", "--games", str(games), "--policy", "xavier"],
            "single_ckpt2.1M_greedy": ["single", "--games", str(games), "--policy", "ckpt"],
            "workers_1thread": ["workers", "--procs", str(cores), "--seconds", str(secs), "--threads", "1"],
            "workers_default_threads": ["workers", "--procs", str(cores), "--seconds", str(secs), "--threads", "default"]}
    out = {"protocol": f"BASELINE.md 3.1-3.2 ({mode}: {games} games per single-process leg, {secs} s per all-cores leg from the first episode's arrival)",
           "host_cores": cores}
    for name, argv in legs.items():
        try:
            r = subprocess.run([sys.executable, script] + argv, capture_output=True, text=True, timeout=secs * 6 + 600)
            lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
            out[name] = json.loads(lines[-1]) if lines else {"unavailable": (r.stderr or "no output")[-300:]}
        except Exception as e:  # noqa: BLE001
            out[name] = {"unavailable": repr(e)[:300]}
        if "unavailable" in out[name] and "no copy of the reference" in str(out[name]["unavailable"]):
            return {"python_reference": None, "reason": out[name]["unavailable"]}
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's algorithm on the host cores (the C oracle port, OpenMP over items).  The Python
    reference itself cannot travel to the GPU box; oracle/bg_oracle.c is its line-by-line restatement, pinned against it."""
    if rank != 0:
        return
    import numpy as np

    from oracle import pyoracle as po

    po.build()
    cores = os.cpu_count() or 1
    n_pos = args.ref_positions
    fx = fixture_positions()
    if fx is not None and n_pos <= len(fx[0]):  # the head of the position set the GPU arm runs on
        boards, players = fx[0][:n_pos], fx[1][:n_pos]
        pos_src = "the first %d positions of the benchmark's set (tests/golden/bench_positions.npz)" % n_pos
    else:
        boards, players = po.random_positions(n_pos, seed=2026)
        pos_src = "%d oracle-generated random-playout positions" % n_pos
    ib, ip, ir = po.all_rolls_items(boards, players)
    packed = packed_random_weights(0).numpy()
    for _ in range(args.warmup):
        po.movegen_eval_bench(ib[: len(ib) // 8], ip[: len(ib) // 8], ir[: len(ib) // 8], packed, H, nthreads=cores)
    t0 = time.perf_counter()
    total = 0
    for _ in range(args.steps):
        n, _ = po.movegen_eval_bench(ib, ip, ir, packed, H, nthreads=cores)
        total += n
    dt = time.perf_counter() - t0
    val = total / dt
    sample = f"{pos_src} x 21 rolls ({len(ib)} items, {total // max(args.steps, 1)} afterstates) per step"
    line = {"impl": "reference", "metric": "afterstates_evaluated_per_sec", "value": val, "unit": "afterstates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / max(args.steps, 1) * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int8 boards / fp32 values", "data": "synthetic",
            "config": {"workload": "config2: random-legal positions x 21 rolls -> legal-move generation + 198-feature encode + 198x128x1 sigmoid-MLP value",
                       "hidden": H, "sample": sample},
            "cpu_baseline": {"value": val, "unit": "afterstates/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "afterstates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--positions", type=int, default=1048576)
    ap.add_argument("--ref-positions", type=int, default=65536)
    ap.add_argument("--cpu-positions", type=int, default=262144)
    ap.add_argument("--selfplay-games", type=int, default=65536)
    ap.add_argument("--selfplay-plies", type=int, default=200)
    ap.add_argument("--selfplay2-plies", type=int, default=20)
    ap.add_argument("--e2e-chunks", type=int, default=0, help="chunks of the host-to-host pipeline (0 = default)")
    ap.add_argument("--e2e-streams", type=int, default=3)
    ap.add_argument("--td0-updates", type=int, default=50)
    ap.add_argument("--cpu-selfplay-games", type=int, default=16384)
    ap.add_argument("--no-selfplay", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sustained", action="store_true", help="skip the ~2 s back-to-back run of the evaluator (roofline.sustained)")
    ap.add_argument("--pyref", default="short", choices=["off", "short", "full"],
                    help="time the UNMODIFIED Python reference when a copy is on the host (baseline/_ref): short = 30 games / 20 s per leg, "
                         "full = BASELINE.md 3.1-3.2 (100 games, 60 s per leg)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
    import mlp_ppo_2ply_multi_b200 as bg

    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- inputs (resident in HBM before the timed region) --------------------------------------------------------
    packed = packed_random_weights(0).to(dev)
    weights = bg.prepare_weights(packed, H)
    boards, players, n_fixture = bench_positions(bg, args.positions, dev, seed=2026 + 7919 * rank, rank=rank)
    P = boards.shape[0]
    B = 21 * P  # items: (position, roll), roll order of src/multi/two_ply.py:10-32
    pool_cap = int(B * 26) + (1 << 20)
    codes = torch.empty(pool_cap, dtype=torch.int64, device=dev)  # compact pool: (code, position) per afterstate
    values = torch.empty(pool_cap, dtype=torch.float32, device=dev)
    ws = torch.empty(bg._lib.lib().bg_movegen_workspace_bytes(B), dtype=torch.uint8, device=dev)

    def step():
        # one pass of the hot path over the batch, one C-ABI call (bg_movegen_eval_all_rolls_compact): position-major move generation into a
        # compact pool + fused encode / value of every afterstate (the evaluator rebuilds each afterstate on chip from its position + move code);
        # the evaluation of the bulk tier's rows overlaps the tail tiers.  Afterstate boards are materialised on demand
        # (bg_afterstates_from_codes), e.g. the one the chosen action leads to; `board_pool` below times the form that writes all of them.
        res, _ = bg.movegen_all_rolls_compact(boards, players, weights, item_cap=500, out_codes=codes, out_values=values, workspace=ws)
        return res

    sampler = ClockSampler(local_rank) if rank == 0 else None  # nvidia-smi needs ~0.5 s to start: begin before the warm-up; idle samples are filtered by power
    for _ in range(max(args.warmup, 3)):
        res = step()
    torch.cuda.synchronize()
    n_after = res.total
    res.raise_for_status()

    # ---- timed region: K steps, CUDA events, barrier + sync on both sides, max over ranks ---------------------------------
    t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    profiling = os.environ.get("BG_PROFILE") == "1"  # ncu --profile-from-start off: capture only the timed region
    if profiling:
        torch.cuda.profiler.start()
    t_beg.record()
    for k in range(args.steps):
        step()
    t_end.record()
    barrier()
    if profiling:
        torch.cuda.profiler.stop()
    total_ms_fused = t_beg.elapsed_time(t_end)
    # per-kernel durations for the roofline: the same pass with the two operators issued back to back on one stream, so that CUDA
    # events on that stream bracket each of them
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * args.steps + 1)]
    ev[0].record()
    for k in range(args.steps):
        r = bg.movegen_all_rolls_compact(boards, players, None, item_cap=500, out_codes=codes, workspace=ws)
        ev[3 * k + 1].record()
        bg.evaluate_codes(r, weights, out=values)
        ev[3 * k + 2].record()
        ev[3 * k + 3].record()
    barrier()
    clocks = sampler.stop() if sampler else None
    # the evaluator alone, back to back for ~2 s: the power-limited (sustained) regime, next to its clocks
    eval_sustained = None
    if rank == 0 and not args.no_sustained:
        smp2 = ClockSampler(local_rank)
        time.sleep(0.6)
        reps = max(10, int(2000.0 / max(ev[1].elapsed_time(ev[2]), 1.0)))
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(reps):
            bg.evaluate_codes(r, weights, out=values)
        s1.record()
        torch.cuda.synchronize()
        ms = s0.elapsed_time(s1) / reps
        c2 = smp2.stop()
        tf = n_after * 2 * 2 * 208 * H / (ms * 1e-3) / 1e12
        eval_sustained = {"what": "k_eval_tc alone, %d launches back to back" % reps, "ms_per_launch": ms, "achieved": tf, "unit": "TFLOP/s", "clocks": c2}
        pk0 = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk0):
            ps = float(json.load(open(pk0)).get("bf16_tflops_sustained", 0.0) or 0.0)
            if ps:
                eval_sustained.update(peak=ps, frac=tf / ps, peak_source="measured (MEASURED_PEAKS.json bf16_tflops_sustained)")
    barrier()
    total_ms = total_ms_fused
    unfused_ms_per_step = ev[0].elapsed_time(ev[3 * args.steps]) / args.steps
    # median over the steps: robust against a single perturbed launch (the value itself is the mean over the fused timed region)
    t_movegen = statistics.median(ev[3 * k].elapsed_time(ev[3 * k + 1]) for k in range(args.steps))
    t_eval = statistics.median(ev[3 * k + 1].elapsed_time(ev[3 * k + 2]) for k in range(args.steps))
    t_movegen_all = [round(ev[3 * k].elapsed_time(ev[3 * k + 1]), 2) for k in range(args.steps)]
    t = torch.tensor([total_ms, float(n_after)], dtype=torch.float64, device=dev)
    if dist is not None:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, n_after_all = float(tmax[0]), float(tsum[1])
    else:
        n_after_all = float(n_after)
    ms_per_step = total_ms / args.steps
    value = n_after_all / (ms_per_step * 1e-3)

    # ---- e2e: public API from HOST buffers (pinned), H2D of the step's inputs and D2H of its result inside the timed region ----
    h_b, h_p = boards.cpu().pin_memory(), players.cpu().pin_memory()
    h_act = torch.empty(B, dtype=torch.int32).pin_memory()
    h_cnt = torch.empty(B, dtype=torch.int32).pin_memory()
    res = r = None
    # ---- the same step with every afterstate board materialised in HBM (bg_movegen_eval_all_rolls: the 52-byte board pool of round 1) ----
    pool = torch.empty((pool_cap, 52), dtype=torch.int8, device=dev)
    pflags = torch.empty(pool_cap, dtype=torch.uint8, device=dev)
    for _ in range(2):
        bg.movegen_evaluate_all_rolls(boards, players, weights, pool, pflags, values, workspace=ws, item_cap=500)
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record()
    for _ in range(args.steps):
        bg.movegen_evaluate_all_rolls(boards, players, weights, pool, pflags, values, workspace=ws, item_cap=500)
    b1.record()
    torch.cuda.synchronize()
    board_ms = b0.elapsed_time(b1) / args.steps
    board_pool = {"what": "bg_movegen_eval_all_rolls: the same step writing every afterstate as a 52-byte board (+ flag) to HBM and reading it back in the evaluator",
                  "ms_per_step": board_ms, "afterstates_per_sec_this_gpu": n_after / (board_ms * 1e-3)}
    # ---- the materialising encoder alone (bg_encode: 52 + 1 bytes in, 198 fp32 out per row): the one purely bandwidth-bound kernel ----
    n_enc = int(min(n_after, 1 << 22))
    feat = torch.empty((n_enc, 198), dtype=torch.float32, device=dev)
    enc_lib = bg._lib.lib()
    def enc():
        bg._lib.check(enc_lib.bg_encode(pool.data_ptr(), pflags.data_ptr(), n_enc, feat.data_ptr(), torch.cuda.current_stream().cuda_stream))
    for _ in range(3):
        enc()
    q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    q0.record()
    for _ in range(10):
        enc()
    q1.record()
    torch.cuda.synchronize()
    t_enc = q0.elapsed_time(q1) / 10
    del feat
    del pool, values, pflags, ws, codes
    torch.cuda.empty_cache()
    n_chunks = args.e2e_chunks if args.e2e_chunks > 0 else (12 if P >= (1 << 19) else 3)
    pipe = bg.HostPipeline(weights, items_per_chunk=(P + n_chunks - 1) // n_chunks, device=dev, item_cap=500, n_streams=args.e2e_streams, all_rolls=True)

    def e2e_step():
        pipe.run(h_b, h_p, None, h_act, h_cnt, temperature=0.0)  # bg_hostpipe_run: chunked inside the library, copies overlap the kernels

    e2e_step()
    barrier()
    pipe.raise_for_status()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    te = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = n_after_all / (float(te[0]) / args.steps * 1e-3)
    h2d = h_b.numel() + h_p.numel()
    d2h = h_act.numel() * 4 + h_cnt.numel() * 4
    n_step_kernels = 5  # k_movegen21<Std> (bulk) + k_movegen21<Big> + k_movegen<4096> (tail tiers) + k_eval_tc twice (bulk-tier rows on the side stream, tail rows after)
    n_e2e_kernels = (n_step_kernels + 2) * n_chunks  # + status fold + k_select, per chunk
    e2e_check = int((h_cnt.to(torch.int64).clamp(max=500)).sum().item())  # must reproduce the afterstate count of the resident path
    pipe.close()
    del pipe

    # ---- roofline of the dominant kernel (algorithmic bytes / measured kernel time) -----------------------------------------
    peak, peak_src = load_peaks()
    eval_bytes = n_after * (8 + 4) + P * (52 + 1)  # (code, position) in + value out per afterstate; position boards / players in
    movegen_bytes = P * (52 + 1) + B * (8 + 4) + n_after * 8  # position in, item offset / count out, (code, position) out per afterstate
    k_mg, k_ev = "bg::k_movegen21<Std> (+ tail tiers k_movegen21<Big>, k_movegen<4096>)", "bg::k_eval_tc (tcgen05, H=128)"
    kern = {k_ev: (t_eval, eval_bytes), k_mg: (t_movegen, movegen_bytes)}
    dom = max(kern, key=lambda k: kern[k][0])
    ach = kern[dom][1] / (kern[dom][0] * 1e-3) / 1e9
    # dram__bytes_read + dram__bytes_write per launch, from the `ncu --set full` capture of this command that scripts/ncu_traffic.py
    # summarised into profiles/r02_traffic.json (valid for the configuration named in that file only)
    traffic, traffic_src = {}, None
    tj = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tj):
        td = json.load(open(tj))
        if int(td.get("positions", -1)) == P:
            traffic = {k_mg: td.get("movegen_bytes"), k_ev: td.get("eval_bytes")}
            traffic_src = "profiles/r02_traffic.json <- " + str(td.get("source"))
    roofline_dom_hbm = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic.get(dom),
                        "traffic_source": traffic_src, "algorithmic_bytes": kern[dom][1], "peak_source": peak_src, "ms_per_launch": kern[dom][0]}
    roofline_movegen = {"bound": "hbm", "kernel": k_mg, "achieved": movegen_bytes / (t_movegen * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": movegen_bytes / (t_movegen * 1e-3) / 1e9 / peak, "traffic": traffic.get(k_mg), "algorithmic_bytes": movegen_bytes,
                        "ms_per_launch": t_movegen, "items_per_sec": B / (t_movegen * 1e-3)}
    enc_bytes = n_enc * (52 + 1 + 198 * 4)
    roofline_encode = {"bound": "hbm", "kernel": "bg::k_encode (materialised 198 fp32 features, parity / hand-off path)", "achieved": enc_bytes / (t_enc * 1e-3) / 1e9,
                       "peak": peak, "unit": "GB/s", "frac": enc_bytes / (t_enc * 1e-3) / 1e9 / peak, "traffic": None, "algorithmic_bytes": enc_bytes,
                       "ms_per_launch": t_enc, "rows": n_enc}
    # the evaluator is a tensor-core kernel: 2 fp16 pieces x (2 * 208 * 128) FLOP per afterstate actually issued to tcgen05
    tc_flops = n_after * 2 * 2 * 208 * H
    # Which measured cuBLAS bf16 figure is the denominator: the BURST one when this run's SM clock (sampled over the timed region) stayed near
    # its maximum -- a kernel "timed alone" --, the SUSTAINED (power-limited) one otherwise.  The evaluator draws ~1 kW: five 54 ms steps do not
    # reach the sustained regime, two seconds of back-to-back launches do -- `sustained` below measures that regime explicitly.
    t_burst = t_sust = None
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        pkd = json.load(open(pk))
        t_burst = float(pkd.get("bf16_tflops", 0.0)) or None
        t_sust = float(pkd.get("bf16_tflops_sustained", 0.0)) or None
    tach = tc_flops / (t_eval * 1e-3) / 1e12
    near_max = bool(clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz") and clocks["sm_mhz"] >= 0.9 * clocks["sm_max_mhz"])
    if t_burst and t_sust:
        use_burst = near_max or tach > t_sust  # a rate above the sustained figure cannot come from the power-limited regime
        tpeak = t_burst if use_burst else t_sust
        tsrc = ("measured (MEASURED_PEAKS.json bf16_tflops, burst: SM clock %s of %s MHz over the timed region)" if use_burst else
                "measured (MEASURED_PEAKS.json bf16_tflops_sustained: SM clock %s of %s MHz over the timed region, power-limited)") % (
                    clocks.get("sm_mhz") if clocks else None, clocks.get("sm_max_mhz") if clocks else None)
    else:
        tpeak, tsrc = 1590.0, "fallback (B200_PROFILING.md 1.59 PFLOP/s)"
    roofline_eval = {"bound": "tensor", "kernel": "bg::k_eval_tc (tcgen05, H=128)", "achieved": tach, "peak": tpeak, "unit": "TFLOP/s",
                     "frac": tach / tpeak, "traffic": traffic.get("bg::k_eval_tc (tcgen05, H=128)"), "peak_source": tsrc, "ms_per_launch": t_eval,
                     "peak_burst": t_burst, "peak_sustained": t_sust, "sustained": eval_sustained,
                     "note": "16-bit tensor FLOPs issued: 2 fp16 weight pieces x 2*208*128 per afterstate (fp32-exact layer 1: one fp16 piece misses the 1e-5 contract); dense fp32-equivalent is 1/2.09 of this. 26 M128 N128 K16 tcgen05.mma per 128-row tile; ncu: tensor pipe 73-75 % active at 1.9 GHz (profiles/r02_ncu_eval_tc_roles.txt); back to back the kernel is POWER limited (~1 kW, SM clock ~1.68 GHz) at the measured sustained cuBLAS bf16 rate"}

    roofline_eval["traffic_source"] = traffic_src
    roofline_eval["kernels_ms"] = {k: v[0] for k, v in kern.items()}
    roofline_eval["movegen_ms_per_step"] = t_movegen_all
    roofline_eval["eval_fp32_tflops_dense_equiv"] = n_after * FLOP_PER_AFTERSTATE / (t_eval * 1e-3) / 1e12
    roofline_eval["hbm_view"] = roofline_dom_hbm if dom == k_ev else None
    # the dominant kernel of the step is the tcgen05 evaluator (tensor-bound); the move generator's roofline is reported next to it
    roofline = roofline_eval if dom == k_ev else dict(roofline_dom_hbm, kernels_ms={k: v[0] for k, v in kern.items()})
    torch.cuda.empty_cache()

    # ---- secondary: BASELINE configs[2] / configs[3], self-play with 65,536 concurrent games per GPU (all ranks, sharded by game id) ----
    def run_selfplay(G, plies, lookahead, la=None, label=""):
        n_local, base = G, rank * G  # weak scaling: G games per GPU, global game ids keep the Philox streams disjoint
        ar = bg.Arena(n_local, hidden_size=H, device=dev, seed=0, game_id_base=base, ring_experiences=n_local * 48, ring_episodes=n_local)
        ar.set_weights(packed, version=1)  # T = 1.5 (reference schedule at version 1)
        if la is not None:
            ar.set_lookahead(*la)
        ar.reset()
        ar.step(120)  # desynchronise game phases (1-ply)
        ar.drain(max_episodes=n_local, max_experiences=n_local * 48)
        barrier()
        s0 = ar.stats()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        done = 0
        chunk = 20 if lookahead == 1 else 5
        while done < plies:
            ar.step(chunk, lookahead=lookahead)
            ar.drain(max_episodes=n_local, max_experiences=n_local * 48)
            done += chunk
        a1.record()
        barrier()
        s1 = ar.stats()
        ar.close()
        keys = ["games", "steps", "passes", "afterstates", "replies", "wait_steps", "errors"]
        d = torch.tensor([float(s1[k] - s0[k]) for k in keys] + [a0.elapsed_time(a1)], dtype=torch.float64, device=dev)
        if dist is not None:
            tot = d.clone()
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            mx = d.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            d = torch.cat([tot[:-1], mx[-1:]])
        v = dict(zip(keys, d[:-1].tolist()))
        ms = float(d[-1])
        return {"workload": label, "games_per_sec": v["games"] / (ms * 1e-3), "afterstates_per_sec": v["afterstates"] / (ms * 1e-3),
                "reply_evals_per_sec": v["replies"] / (ms * 1e-3), "plies_per_sec": world * G * done / (ms * 1e-3), "ms_per_ply_step": ms / done,
                "games_finished": int(v["games"]), "mean_steps_per_game": v["steps"] / max(v["games"], 1.0),
                "pass_rate": v["passes"] / max(v["steps"], 1.0), "wait_steps": int(v["wait_steps"]), "errors": int(v["errors"])}

    selfplay = selfplay2 = selfplay2b = selfplay2s = None
    if not args.no_selfplay:
        G = args.selfplay_games
        selfplay = run_selfplay(G, args.selfplay_plies, 1, None,
                                f"config3: {G} concurrent 1-ply self-play games per GPU, T=1.5, Philox dice, episodes drained on device")
        selfplay2 = run_selfplay(G, args.selfplay2_plies, 2, (4, 5, 1.0, 0.9),
                                 f"config4 (reference setting, two_ply.py): {G} games per GPU, top-4 candidates x 21 rolls, mean of top-5 replies, score = S - 0.9 W")
        # the lookahead exactly as shipped: the reference evaluates only random.sample(replies, 50) of the rolls 1-1 / 2-2 / 3-3
        # (two_ply.py:119-121); here a reproducible Philox-keyed sample taken before evaluation (bg_two_ply_reply_sampling)
        bg.set_reply_sampling(50, 2026)
        try:
            selfplay2s = run_selfplay(G, args.selfplay2_plies, 2, (4, 5, 1.0, 0.9),
                                      f"config4 as shipped (two_ply.py incl. random.sample(replies, 50) on 1-1 / 2-2 / 3-3): {G} games per GPU, top-4 candidates x 21 rolls, mean of top-5 replies")
        finally:
            bg.set_reply_sampling(0, 0)
        G2 = max(G // 8, 1)
        selfplay2b = run_selfplay(G2, args.selfplay2_plies, 2, (0, 1, 1.0, 1.0),
                                  f"config4 (north-star expectimax): {G2} games per GPU, ALL candidates x 21 rolls, best reply, score = S - W")

    # ---- BASELINE configs[4]: full TD(0) loop -- arena self-play -> 200-episode Trainer.update -> weights published (one broadcast) ----
    def run_td0_loop(G, n_updates):
        from mlp_ppo_2ply_multi_b200 import distributed as bgd

        pm = bg.ParameterManager(hidden_size=H)
        ar = bg.Arena(G, hidden_size=H, device=dev, seed=1, game_id_base=rank * G, ring_experiences=G * 48, ring_episodes=G)
        pm.subscribe(ar)
        if rank == 0:
            pm.set_packed(packed, H)  # collective when world > 1
            tr = bg.Trainer(pm, device=dev)
        else:
            pm.sync_from_source()
        ar.reset()
        ar.step(120)
        ar.drain(max_episodes=G, max_experiences=G * 48)

        lstream = torch.cuda.Stream(device=dev)  # the learner kernel (one 8-CTA cluster) overlaps the next self-play ply
        n_iter = [0]

        staged = [None]  # (batch, event) gathered in the previous iteration
        unpublished = [False]  # an update launched by rank 0 that has not been published yet (tracked identically on every rank)
        host = {}  # host time per section of an iteration (perf_counter; the drains include their wait for the GPU)

        trace = [] if os.environ.get("BG_LOOP_TRACE") else None  # development: GPU timeline of the loop (events on both streams)

        def mark(row, name, stream=None):
            if trace is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record(stream if stream is not None else torch.cuda.current_stream(dev))
                row.append((name, e))

        def lap(name, t0):
            t1 = time.perf_counter()
            host[name] = host.get(name, 0.0) + (t1 - t0)
            return t1

        def iteration(timed):
            # Host order matters: the loop is paced by what the HOST has to do between two learner launches.  (1) publish update u-1 and
            # launch update u on the batch staged in the previous iteration -- nothing the host did not already have; (2) the next ply of
            # every game (|| update u on its own stream), drain 200 // world finished episodes, gather them for update u+1, drop the
            # surplus (the sequential learner is the bottleneck).  The drains read counts back: they block the host behind everything
            # enqueued on the arena's stream -- including the wait for update u-1 -- which is why they come AFTER the learner launch.
            t = time.perf_counter()
            m = None
            row = []
            if trace is not None:
                trace.append(row)
            if rank == 0:
                mark(row, "main:top")
                m = tr.finish(metrics=False)  # stream-ordered wait, set_packed -> (broadcast) -> arena.set_weights; no host read-back
                mark(row, "main:published")
                t = lap("publish", t)
                if staged[0] is not None:
                    # the update needs its batch, not the publication enqueued above (the learner owns its weights; what is published
                    # is a snapshot taken on the learner's stream)
                    lstream.wait_event(staged[0][1])
                    with torch.cuda.stream(lstream):
                        if timed is not None:
                            timed[0].record()
                        mark(row, "learner:start")
                        tr.update_async(staged[0][0])
                        mark(row, "learner:end")
                        if timed is not None:
                            timed[1].record()
                t = lap("launch_update", t)
            elif unpublished[0]:  # rank 0 publishes the update it launched in the previous iteration
                pm.sync_from_source()
                t = lap("publish", t)
            unpublished[0] = staged[0] is not None  # (the same on every rank) rank 0 launches an update in this iteration
            ar.step(1)
            mark(row, "main:ply done")
            t = lap("step", t)
            quota = 200 // world  # every rank's arena feeds the trainer, as every worker process feeds the reference's queue
            batch = ar.drain(max_episodes=quota)
            while batch.n_episodes < quota:  # not reached with tens of thousands of games in flight
                ar.step(1)
                batch = ar.drain(max_episodes=quota)
            t = lap("drain_quota(blocks)", t)
            if world > 1:
                # each field the trainer reads gathered straight into its final padded array (two buffer sets in turn: update u may still
                # be reading the previous one): no host sync, no unpacking copies, one coalesced NCCL launch
                batch = bgd.all_gather_episodes(batch, quota, quota * 300, compact=False, fields=bgd.LEARNER_FIELDS)
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(dev))
            staged[0] = (batch, ready)
            n_iter[0] += 1
            t = lap("gather", t)
            ar.drain(max_episodes=G, max_experiences=G * 48)
            mark(row, "main:drained")
            lap("drain_surplus(blocks)", t)
            return m

        def flush():  # publish what is in flight and read its metrics (every rank takes part in the broadcast)
            m = None
            if rank == 0:
                m = tr.finish()
            elif unpublished[0]:
                pm.sync_from_source()
            unpublished[0] = False
            return m

        for _ in range(3):
            iteration(None)
        flush()  # warm-up includes one metrics read-back: its ~25 small device ops load their modules outside the timed region
        barrier()
        s0 = ar.stats()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_updates)]
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        last = None
        host.clear()
        h0 = time.perf_counter()
        for u in range(n_updates):
            last = iteration(evs[u]) or last
        host_ms = (time.perf_counter() - h0) * 1e3 / max(n_updates, 1)  # enqueue time per iteration (no synchronisation inside the loop)
        last = flush() or last
        a1.record()
        barrier()
        if trace and rank == 0:
            torch.cuda.synchronize()
            base = trace[-12][0][1]
            for row in trace[-12:-4]:
                print("trace", [(n, round(base.elapsed_time(e), 3)) for n, e in row], file=sys.stderr)
        s1 = ar.stats()
        ms = a0.elapsed_time(a1)
        d = torch.tensor([float(s1["games"] - s0["games"]), float(s1["wait_steps"] - s0["wait_steps"]), ms], dtype=torch.float64, device=dev)
        if dist is not None:
            tot = d.clone()
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            mx = d.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            d = torch.cat([tot[:2], mx[2:]])
        out = None
        if rank == 0:
            upd_ms = sum(x.elapsed_time(y) for x, y in evs) / n_updates
            ms = float(d[2])
            out = {"workload": f"config5: {G} self-play games per GPU -> Trainer.update on 200-episode batches (sequential TD(0)/Adam, rank 0) -> packed weights "
                               f"published to every arena ({'episodes all-gathered from every rank, one NCCL broadcast of the weights' if world > 1 else 'single GPU'}); surplus episodes dropped",
                   "updates_per_sec": n_updates / (ms * 1e-3), "episodes_trained_per_sec": 200 * n_updates / (ms * 1e-3),
                   "games_played_per_sec": float(d[0]) / (ms * 1e-3), "ms_per_update_kernel": upd_ms, "ms_per_iteration": ms / n_updates, "host_enqueue_ms_per_iteration": host_ms, "host_ms_per_section": {k: round(v * 1e3 / max(n_updates, 1), 4) for k, v in host.items()},
                   "actor_wait_steps": int(d[1]), "weights_version": pm.get_version(), "temperature": pm.get_temperature(),
                   "last_update": {k: v for k, v in last.items() if isinstance(v, float)}}
        ar.close()
        return out

    # learner alone: one 200-episode update on a fixed drained batch (kernel time, CUDA events)
    def run_learner():
        if rank != 0:
            return None
        ar = bg.Arena(8192, hidden_size=H, device=dev, seed=2)
        ar.set_weights(packed, version=1)
        ar.reset()
        # steady-state episodes: the FIRST games to finish in a fresh arena are the shortest ones (41 experiences per episode instead of ~90,
        # which made this line 1.09 ms where the training loop's updates take 1.5 ms), so play 400 plies (every game has restarted a few
        # times), drop what finished, and take the next 200 episodes
        ar.step(400)
        ar.drain(max_episodes=8192, max_experiences=8192 * 48)
        batch = ar.drain(max_episodes=200)
        while batch.n_episodes < 200:
            ar.step(4)
            batch = ar.drain(max_episodes=200)
        ar.close()
        L = bg.TD0Learner(H, dev)
        L.set_parameters(packed, reset_optimizer=True)
        for _ in range(3):
            L.update_batch(batch, check_status=False)
        torch.cuda.synchronize()
        R = 10
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(R):
            L.update_batch(batch, check_status=False)
        a1.record()
        torch.cuda.synchronize()
        ms = a0.elapsed_time(a1) / R
        out = {"workload": "Trainer.update on one 200-episode batch drained from the arena (sequential per-episode forward, TD(0) targets, backward, "
                           "clip_grad_norm_, Adam), one kernel launch (k_td0_update_tc, 8-CTA cluster)",
               "ms_per_update": ms, "episodes_per_sec": 200 / (ms * 1e-3), "experiences_per_sec": batch.n_experiences / (ms * 1e-3),
               "us_per_optimizer_step": ms * 1e3 / 200, "experiences": int(batch.n_experiences)}
        if not args.no_cpu_baseline and world == 1:
            from oracle import pyoracle as po

            po.build()
            ob, of = batch.observation_boards()
            N = batch.n_experiences
            O = po.Learner(packed.cpu().numpy(), H)
            t0 = time.perf_counter()
            O.update(ob[:N].cpu().numpy(), of[:N].cpu().numpy(), batch.reward[:N].cpu().numpy(), batch.ep_offsets[:201].cpu().numpy())
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": 200 / dt, "unit": "episodes/s", "cores": 1, "kind": "port",
                                   "sample": f"the same 200 episodes through the C restatement of Trainer.update (inherently sequential), {dt:.2f} s"}
        return out

    learner = td0 = None
    if not args.no_selfplay:
        learner = run_learner()
        td0 = run_td0_loop(args.selfplay_games, args.td0_updates)

    # ---- CPU baseline: the oracle port on the host cores, bounded sample of the same workload -----------------------------------
    cpu = None
    if not args.no_cpu_baseline and rank == 0 and world == 1:  # reported baseline, rank 0 at N=1 only
        from oracle import pyoracle as po

        po.build()
        cores = os.cpu_count() or 1
        ns = min(args.cpu_positions, boards.shape[0])
        sb, sp, sr = po.all_rolls_items(boards[:ns].cpu().numpy(), players[:ns].cpu().numpy())
        pk = packed.cpu().numpy()
        po.movegen_eval_bench(sb[:4096], sp[:4096], sr[:4096], pk, H, nthreads=cores)
        t0 = time.perf_counter()
        n_cpu, _ = po.movegen_eval_bench(sb, sp, sr, pk, H, nthreads=cores)
        dt = time.perf_counter() - t0
        cpu = {"value": n_cpu / dt, "unit": "afterstates/s", "cores": cores, "kind": "port",
               "sample": f"first {ns} of the benchmark's positions x 21 rolls ({len(sb)} items, {n_cpu} afterstates), {dt:.1f} s, OpenMP"}
        if selfplay is not None:  # BASELINE configs[0]: the reference's CPU self-play loop (1-ply, T = 1.5), one game per thread at a time
            ng = args.cpu_selfplay_games
            t0 = time.perf_counter()
            n_after_sp, n_steps, n_dec = po.selfplay_bench(pk, H, 1.5, ng, seed=0, nthreads=cores)
            dt = time.perf_counter() - t0
            selfplay["cpu_baseline"] = {"value": ng / dt, "unit": "games/s", "cores": cores, "kind": "port", "afterstates_per_sec": n_after_sp / dt,
                                        "sample": f"{ng} complete 1-ply self-play games (T=1.5) through the C restatement of Worker.play_episode, "
                                                  f"{n_steps / ng:.1f} plies/game, {dt:.1f} s, OpenMP over games"}

        # ---- the UNMODIFIED Python reference on the same host cores (BASELINE.md 3.1-3.2, configs[0]), when a copy is on the host ----
        cpu["python_reference"] = run_python_reference(args.pyref, cores)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    line = {"metric": "afterstates_evaluated_per_sec", "value": value, "unit": "afterstates/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int8 boards / fp32 values", "data": "synthetic",
            "config": {"workload": "config2: random-legal positions x 21 rolls -> legal-move generation + fused 198-feature encode + 198x128x1 sigmoid-MLP value",
                       "positions_per_gpu": int(boards.shape[0]), "items_per_gpu": int(B), "afterstates_per_gpu_step": int(n_after), "hidden": H,
                       "positions_from_fixture": int(n_fixture),
                       "l2": "inputs+outputs per step (>25 GB) far exceed the 126 MB L2", "parallelism": f"{world} x independent shards"},
            "e2e": {"value": e2e_value, "unit": "afterstates/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "what": f"bg_hostpipe_run (C ABI; bg.HostPipeline): pinned host positions + players -> {n_chunks} chunks on {args.e2e_streams} library streams (H2D, bg_movegen_eval_all_rolls_compact, bg_select(greedy), D2H) -> 21 host (action, count) pairs per position",
                    "afterstates_check": e2e_check},
            "gpu_launches": n_step_kernels * args.steps,
            "gpu_launches_note": f"timed region, per step (bg_movegen_eval_all_rolls_compact): k_movegen21<Std> (bulk tier) + tail tiers k_movegen21<Big>, k_movegen<4096> + k_eval_tc twice (bulk-tier rows on the side stream, tail-tier rows after); unfused ms_per_step {unfused_ms_per_step:.2f}; the e2e region launches {n_e2e_kernels} per step (the same five + status fold + k_select, per chunk)",
            "board_pool": board_pool, "roofline": roofline, "roofline_movegen": roofline_movegen, "roofline_eval": roofline_eval, "roofline_encode": roofline_encode, "cpu_baseline": cpu, "clocks": clocks, "selfplay_1ply": selfplay, "selfplay_2ply": selfplay2, "selfplay_2ply_sampled_replies": selfplay2s,
            "selfplay_2ply_all_candidates": selfplay2b, "learner": learner, "td0_loop": td0}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
