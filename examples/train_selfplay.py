#!/usr/bin/env python
"""Self-play TD(0) training on one or more B200s: the loop of the reference's src/main.py:84-133 (workers -> ExperienceQueue ->
Trainer.update on 200-episode batches -> ParameterManager.set_parameters -> workers pick up the new weights and temperature) with
the GPU arena as the actor and the cluster learner as the trainer.

    python examples/train_selfplay.py --updates 500                                   # one GPU
    python -m torch.distributed.run --nproc-per-node 8 examples/train_selfplay.py     # every rank's arena feeds rank 0's learner
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlp_ppo_2ply_multi_b200 as bg  # noqa: E402
from mlp_ppo_2ply_multi_b200 import distributed as bgd  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=65536, help="concurrent self-play games per GPU")
    ap.add_argument("--hidden", type=int, default=128)
    ap.add_argument("--updates", type=int, default=200, help="Trainer.update calls (200 episodes each)")
    ap.add_argument("--eval-every", type=int, default=100, help="play a match against the initial weights every N updates (0 = never)")
    ap.add_argument("--save", default="", help="checkpoint file name under ./models (reference .pth format)")
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device(f"cuda:{int(os.environ.get('LOCAL_RANK', '0'))}")
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)

    torch.manual_seed(args.seed)
    pm = bg.ParameterManager(hidden_size=args.hidden)           # reference: Manager().dict() + lock + version
    arena = bg.Arena(args.games, hidden_size=args.hidden, device=dev, seed=args.seed, game_id_base=rank * args.games)
    pm.subscribe(arena)                                         # local: this rank's current weights
    pm.publish()                                                # collective when world > 1: rank 0's initial weights everywhere
    trainer = bg.Trainer(pm, device=dev) if rank == 0 else None
    initial = bg.pack_weights(pm.get_parameters()).to(dev)
    learn_stream = torch.cuda.Stream(device=dev)
    quota = 200 // world
    arena.reset()
    t0, metrics = time.time(), {}
    staged = None                                               # (batch, event) gathered in the previous iteration
    for u in range(args.updates + 1):                           # iteration u launches update u-1 (the batch is staged one iteration ahead)
        # host order: publish + launch first (nothing the host has to wait for), then the ply / drain / gather for the NEXT update --
        # the drains read counts back and would otherwise hold the host behind the wait for the running update
        if rank == 0:
            metrics = trainer.finish() or metrics               # publish update u-2 (one broadcast when world > 1)
            if staged is not None:
                learn_stream.wait_event(staged[1])              # the learner waits for its batch, NOT for the publication enqueued above
                with torch.cuda.stream(learn_stream):
                    trainer.update_async(staged[0])             # 200 sequential TD(0)/Adam steps, one kernel launch, no host sync
        elif u > 1:
            pm.sync_from_source()
        arena.step(1)                                           # next ply of every game || the running update on learn_stream
        batch = arena.drain(max_episodes=quota)
        while batch.n_episodes < quota:
            arena.step(1)
            batch = arena.drain(max_episodes=quota)
        if world > 1:
            batch = bgd.all_gather_episodes(batch, quota, quota * arena.max_plies, compact=False)  # two buffer sets in turn
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(dev))
        staged = (batch, ready)
        arena.drain(max_episodes=args.games, max_experiences=args.games * 48)  # the sequential learner is the bottleneck: drop the surplus
        if rank == 0 and u and args.eval_every and u % args.eval_every == 0:
            # evaluate the weights that ARE published.  Publication is a collective: calling trainer.finish() here would
            # issue a broadcast the other ranks do not take part in
            if world == 1:
                trainer.finish()
            now = bg.pack_weights(pm.get_parameters()).to(dev)
            m = bg.play_match(now, initial, n_games=4096, hidden_size=args.hidden, device=dev, seed=u)
            print(f"update {u:6d}  version {pm.get_version()}  T {pm.get_temperature():.3f}  loss {metrics.get('Loss/Training Loss', 0):.5f}  "
                  f"len {metrics.get('Episode/Average Episode Length', 0):.1f}  vs initial: win {m['a_win_rate']:.3f}  ppg {m['a_points_per_game']:+.3f}  "
                  f"[{u * 200 / (time.time() - t0):,.0f} episodes/s]", flush=True)
    if rank == 0:
        trainer.finish()
        if args.save:
            print("saved", pm.save_model(args.save))
    elif args.updates:
        pm.sync_from_source()
    stats = bgd.all_reduce_stats(arena.stats(), device=dev)
    if rank == 0:
        print({k: stats[k] for k in ("games", "steps", "afterstates", "win_regular", "win_gammon", "win_backgammon")})
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
