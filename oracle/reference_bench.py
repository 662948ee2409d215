#!/usr/bin/env python
"""Times the UNMODIFIED Python reference (BASELINE.json configs[0]) on this host's cores.

TEST / BASELINE INFRASTRUCTURE: bench.py's cpu_baseline leg runs this file as a subprocess when a copy of the reference is present
(BG_REFERENCE_PATH, /root/reference in the build container, baseline/_ref on the GPU box -- __graft_entry__.build() stages it);
the product never imports it.  The reference's own code does all the work, through oracle/reference_shim.py (stubs for gym /
tensorboardX / boto3, which the reference imports but this path does not use):

  single   Worker.play_episode (src/multi/worker.py:78-174), one process, torch.set_num_threads(1), np.random.seed(0),
           torch.manual_seed(0): `xavier` = default BackgammonPolicyNetwork() at T = 1.5, `ckpt` = the shipped
           backgammon_256_standard_episode_2100000.pth at T -> 0 (argmax up to ties).  games/s, decisions/s, afterstates/s
           (= sum of env.num_moves over decisions), plies/game.
  workers  worker_function x C spawned processes with the real ParameterManager (multiprocessing.Manager proxies) and
           ExperienceQueue (src/main.py:63-91), measured from the first episode's arrival for `seconds`; torch threads as shipped
           (`default`) or 1 per worker.

Prints one JSON line.
    python oracle/reference_bench.py single --games 100 --policy xavier
    python oracle/reference_bench.py workers --procs 16 --seconds 60 --threads 1
"""
import argparse
import contextlib
import io
import json
import os
import queue
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


if os.environ.get("BG_REFBENCH_CHILD"):  # a spawned worker re-imports this module before it unpickles its arguments (the reference's
    # ParameterManager / ExperienceQueue instances), so the import shim has to be in place by then
    from oracle import reference_shim as _shim

    _shim.REF = os.environ["BG_REFBENCH_CHILD"]
    _shim.install()


def find_reference():
    for p in (os.environ.get("BG_REFERENCE_PATH"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if p and os.path.isdir(os.path.join(p, "src", "backgammon")):
            return p
    return None


def _worker_entry(worker_id, parameter_manager, experience_queue, ref, one_thread):
    """child process: install the shim, then hand over to the reference's worker_function"""
    os.environ["BG_REFERENCE_PATH"] = ref
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import reference_shim as shim

    shim.REF = ref
    shim.install()
    import torch

    if one_thread:
        torch.set_num_threads(1)
    from multi.worker import worker_function

    with contextlib.redirect_stdout(io.StringIO()):
        worker_function(worker_id, parameter_manager, experience_queue)


def run_single(ref, games, policy):
    from oracle import reference_shim as shim

    shim.REF = ref
    shim.install()
    import numpy as np
    import torch

    torch.set_num_threads(1)
    from agents import BackgammonPolicyNetwork
    from environments import BackgammonEnv
    from multi.worker import Worker

    np.random.seed(0)
    torch.manual_seed(0)
    net = BackgammonPolicyNetwork()
    temperature = 1.5
    if policy == "ckpt":
        net.load_state_dict(torch.load(os.path.join(ref, "src", "play", "backgammon_256_standard_episode_2100000.pth"), map_location="cpu"))
        temperature = 1e-4  # softmax(V / T) collapses onto the argmax (play_versus_ai.py:188-195 is the greedy rule)
    w = Worker.__new__(Worker)  # __init__ only wires the ParameterManager; play_episode is the unmodified hot loop
    w.worker_id, w.device, w.policy_network, w.temperature = 0, torch.device("cpu"), net, temperature

    class CountingEnv(BackgammonEnv):
        afterstates = 0
        steps = 0

        def step(self, action):
            if action is not None:
                CountingEnv.afterstates += self.num_moves
            CountingEnv.steps += 1
            return super().step(action)

    env = CountingEnv()
    decisions = 0
    t0 = time.perf_counter()
    for _ in range(games):
        with contextlib.redirect_stdout(io.StringIO()):
            ep = w.play_episode(env)
        decisions += len(ep.experiences)
    dt = time.perf_counter() - t0
    return {"mode": "single", "policy": policy, "temperature": temperature, "games": games, "seconds": dt, "games_per_sec": games / dt,
            "decisions_per_sec": decisions / dt, "afterstates_per_sec": CountingEnv.afterstates / dt, "plies_per_game": CountingEnv.steps / games,
            "processes": 1, "torch_threads": 1}


def run_workers(ref, procs, seconds, threads):
    import multiprocessing as mp

    from oracle import reference_shim as shim

    shim.REF = ref
    shim.install()
    from multi.experience_queue import ExperienceQueue
    from multi.parameter_manager import ParameterManager

    os.environ["BG_REFBENCH_CHILD"] = ref
    mp.set_start_method("spawn", force=True)  # as src/main.py:164 does, before any multiprocessing object exists
    ctx = mp
    manager = ctx.Manager()
    pm = ParameterManager(manager.Lock(), manager.Value("i", 1), manager.dict())
    q = ExperienceQueue()
    ps = [ctx.Process(target=_worker_entry, args=(i, pm, q, ref, threads == "1"), daemon=True) for i in range(procs)]
    for p in ps:
        p.start()
    n, decisions, t_first = 0, 0, None
    deadline = None
    t_start = time.perf_counter()
    while True:
        try:
            ep = q.get(timeout=1)
        except queue.Empty:
            if deadline is not None and time.perf_counter() >= deadline:
                break
            if deadline is None and (time.perf_counter() - t_start > 180 or not any(p.is_alive() for p in ps)):
                for p in ps:
                    p.terminate()
                manager.shutdown()
                return {"mode": "workers", "processes": procs, "torch_threads": threads, "unavailable": "no episode arrived within 180 s (workers dead?)"}
            continue
        now = time.perf_counter()
        if t_first is None:  # the clock starts when the first episode arrives (process start-up and imports excluded)
            t_first, deadline = now, now + seconds
            continue
        if now >= deadline:
            break
        n += 1
        decisions += len(ep.experiences)
    dt = time.perf_counter() - t_first
    for p in ps:
        p.terminate()
    for p in ps:
        p.join(timeout=5)
    manager.shutdown()
    return {"mode": "workers", "processes": procs, "torch_threads": threads, "seconds": dt, "games": n, "games_per_sec": n / dt,
            "decisions_per_sec": decisions / dt}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["single", "workers"])
    ap.add_argument("--games", type=int, default=100)
    ap.add_argument("--policy", default="xavier", choices=["xavier", "ckpt"])
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--threads", default="1", choices=["default", "1"])
    a = ap.parse_args()
    ref = find_reference()
    if ref is None:
        print(json.dumps({"mode": a.mode, "unavailable": "no copy of the reference on this host (BG_REFERENCE_PATH, /root/reference, baseline/_ref)"}))
        return
    out = run_single(ref, a.games, a.policy) if a.mode == "single" else run_workers(ref, a.procs, a.seconds, a.threads)
    out["host_cores"] = os.cpu_count()
    out["reference"] = ref
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
