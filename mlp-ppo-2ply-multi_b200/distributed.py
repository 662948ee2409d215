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


_EP_FIELDS = (("after_boards", torch.int8, 52), ("meta", torch.uint8, 1), ("reward", torch.float32, 1), ("state_value", torch.float32, 1),
              ("next_state_value", torch.float32, 1), ("n_moves", torch.int16, 1), ("action", torch.int16, 1), ("roll", torch.uint8, 2))


def all_gather_episodes(batch, max_episodes: int, max_experiences: int, group=None, compact: bool = True):
    """Config 5 (collective 3): every rank contributes its drained episodes (at most max_episodes /
    max_experiences) and receives all of them in rank order as one EpisodeBatch -- what the trainer rank feeds to Trainer.update.
    ONE all_gather of a fixed-size byte buffer per rank (compact records: 72 B per experience, so 200 episodes are ~1.3 MB in
    total); replaces the reference's ExperienceQueue.put from every worker process (src/multi/worker.py:60-64).

    compact=True : the result is a dense CSR batch (one small host read-back for the per-rank sizes).
    compact=False: no host synchronisation at all -- every rank's records stay in their padded segment and the batch carries explicit
                   episode lengths (EpisodeBatch.ep_len; bg_learner_update's ep_len argument); n_episodes = world * max_episodes, ranks
                   that supplied fewer episodes contribute zero-length ones, which the learner skips."""
    from .episode import EpisodeBatch

    E, N = int(batch.n_episodes), int(batch.n_experiences)
    if E > max_episodes or N > max_experiences:
        raise ValueError(f"batch ({E} episodes, {N} experiences) exceeds the gather quota ({max_episodes}, {max_experiences})")
    dev = batch.after_boards.device
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1

    def pad_bytes(t, rows, cols, dtype, used):
        buf = torch.zeros((rows, cols), dtype=dtype, device=dev)
        if used:
            buf[:used] = t[:used].reshape(used, cols)
        return buf.reshape(-1).view(torch.uint8)

    parts = [torch.tensor([E, N], dtype=torch.int64, device=dev).view(torch.uint8)]
    parts += [pad_bytes(getattr(batch, name), max_experiences, cols, dt, N) for name, dt, cols in _EP_FIELDS]
    parts.append(pad_bytes(batch.ep_offsets[: E + 1] if E else batch.ep_offsets[:1], max_episodes + 1, 1, torch.int64, E + 1))
    cols_info = batch.ep_info.shape[1]
    parts.append(pad_bytes(batch.ep_info, max_episodes, cols_info, torch.int32, E))
    mine = torch.cat(parts)
    if world > 1:
        allb = torch.empty((world, mine.numel()), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allb, mine, group=group) if dev.type == "cuda" else dist.all_gather(list(allb.unbind(0)), mine, group=group)
    else:
        allb = mine.reshape(1, -1)

    def field(pos, rows, cols, dt):  # [world, rows, cols] view of one field of every rank (one strided copy, freshly aligned)
        nb = rows * cols * torch.empty(0, dtype=dt).element_size()
        return allb[:, pos:pos + nb].clone(memory_format=torch.contiguous_format).view(dt).reshape(world, rows, cols), pos + nb  # clone: fresh, aligned, dense storage

    hdr, pos = allb[:, :16].clone(memory_format=torch.contiguous_format).view(torch.int64).reshape(world, 2), 16
    f = {}
    for name, dt, cols in _EP_FIELDS:
        f[name], pos = field(pos, max_experiences, cols, dt)
    offs, pos = field(pos, max_episodes + 1, 1, torch.int64)
    offs = offs.reshape(world, max_episodes + 1)
    info, pos = field(pos, max_episodes, cols_info, torch.int32)

    def flat(name, sel=None):
        t = f[name] if sel is None else f[name][sel]
        cols = t.shape[-1]
        return t.reshape(-1, cols).contiguous() if cols > 1 else t.reshape(-1).contiguous()

    if not compact:
        j = torch.arange(max_episodes, device=dev).reshape(1, -1)
        ep_len = torch.where(j < hdr[:, 0:1], offs[:, 1:] - offs[:, :-1], torch.zeros_like(offs[:, 1:])).clamp_(min=0)
        begin = offs[:, :-1] + torch.arange(world, device=dev).reshape(-1, 1) * max_experiences
        ep_offsets = torch.cat([begin.reshape(-1), torch.full((1,), world * max_experiences, dtype=torch.int64, device=dev)])
        return EpisodeBatch(world * max_episodes, world * max_experiences, flat("after_boards"), flat("meta"), flat("reward"), flat("state_value"),
                            flat("next_state_value"), flat("n_moves"), flat("action"), flat("roll"), ep_offsets.contiguous(),
                            info.reshape(-1, cols_info).contiguous(), ep_len=ep_len.reshape(-1).to(torch.int32).contiguous())
    sizes = hdr.tolist()  # the one host read-back
    rows = torch.cat([torch.arange(n_r, device=dev) + r * max_experiences for r, (_, n_r) in enumerate(sizes)]) if sizes else None
    o_parts, tot = [torch.zeros(1, dtype=torch.int64, device=dev)], 0
    for r, (e_r, n_r) in enumerate(sizes):
        if e_r:
            o_parts.append(offs[r, 1:e_r + 1] - offs[r, 0] + tot)
        tot += n_r
    eps = torch.cat([torch.arange(e_r, device=dev) + r * max_episodes for r, (e_r, _) in enumerate(sizes)])
    pick = lambda name: flat(name)[rows]  # noqa: E731
    return EpisodeBatch(sum(e for e, _ in sizes), tot, pick("after_boards"), pick("meta"), pick("reward"), pick("state_value"),
                        pick("next_state_value"), pick("n_moves"), pick("action"), pick("roll"), torch.cat(o_parts).contiguous(),
                        info.reshape(-1, cols_info)[eps].contiguous())
