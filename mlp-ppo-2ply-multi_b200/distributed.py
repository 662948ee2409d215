"""Multi-GPU plumbing (one process per GPU, torch.distributed).  Games are independent (reference README.md:59-61; each worker
owns its env, src/multi/worker.py:52), so the arena shards BY GAME with no data-path collective.  The only collectives are the
ones SURVEY.md section 8(e) names: one broadcast of the ~104 KB packed weight blob (+ version, temperature) after each trainer
step, and a sum-reduction of the small statistics vector."""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist

from .arena import STAT_NAMES


def shard_games(n_games_total: int, rank: int, world: int) -> Tuple[int, int]:
    """(n_local, game_id_base): contiguous blocks; per-game Philox streams are keyed by the GLOBAL game id, so a game's dice and
    sampled actions do not depend on how many GPUs the arena is sharded over."""
    base, rem = divmod(n_games_total, world)
    n_local = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return n_local, start


def broadcast_weights(packed: torch.Tensor, version: int, temperature: float, src: int = 0, group=None):
    """One collective: the packed fp32 blob with (version, temperature) appended.  Returns (packed, version, temperature)."""
    blob = torch.cat([packed.reshape(-1).to(torch.float32), torch.tensor([float(version), float(temperature)], device=packed.device)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(blob, src=src, group=group)
    return blob[:-2], int(blob[-2].item()), float(blob[-1].item())


def all_reduce_stats(stats: Dict[str, int], device=None, group=None) -> Dict[str, int]:
    """Sum the arena statistics over ranks (episodes, plies, afterstates, win types ...)."""
    t = torch.tensor([int(stats.get(k, 0)) for k in STAT_NAMES], dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return dict(zip(STAT_NAMES, t.tolist()))
