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
    """One collective: the packed fp32 blob with (version, temperature) appended.  Returns (packed, version, temperature); reads the two
    scalars back on the host (one synchronisation) -- ParameterManager.publish avoids that by deriving them on every rank."""
    blob = torch.cat([packed.reshape(-1).to(torch.float32), torch.tensor([float(version), float(temperature)], device=packed.device)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(blob, src=src, group=group)
    tail = blob[-2:].tolist()
    return blob[:-2], int(tail[0]), float(tail[1])


def all_reduce_stats(stats: Dict[str, int], device=None, group=None) -> Dict[str, int]:
    """Sum the arena statistics over ranks (episodes, plies, afterstates, win types ...)."""
    t = torch.tensor([int(stats.get(k, 0)) for k in STAT_NAMES], dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return dict(zip(STAT_NAMES, t.tolist()))


_EP_FIELDS = (("after_boards", torch.int8, 52), ("meta", torch.uint8, 1), ("reward", torch.float32, 1), ("state_value", torch.float32, 1),
              ("next_state_value", torch.float32, 1), ("n_moves", torch.int16, 1), ("action", torch.int16, 1), ("roll", torch.uint8, 2))


class _GatherWorkspace:
    """per-(device, quota, world) send / receive buffers of all_gather_episodes(compact=False), allocated once.  SLOTS buffer sets are handed
    out in turn, so a returned batch stays valid until the SECOND next call with the same quota: the consumer of batch u (the learner kernel, on
    its own stream) may still be reading it while batch u + 1 is being gathered."""

    SLOTS = 2
    cache: dict = {}
    coalescing = True  # cleared the first time the backend refuses a coalesced all-gather

    @classmethod
    def get(cls, dev, world, max_episodes, max_experiences, cols_info):
        key = (str(dev), world, max_episodes, max_experiences, cols_info)
        ring = cls.cache.get(key)
        if ring is None:
            ring = {"next": 0, "sets": []}
            for _ in range(cls.SLOTS):
                w = {"send": {}, "recv": {}}
                for name, dt, cols in _EP_FIELDS:
                    shape = (max_experiences, cols) if cols > 1 else (max_experiences,)
                    w["send"][name] = torch.zeros(shape, dtype=dt, device=dev)
                    w["recv"][name] = torch.zeros((world * max_experiences,) + shape[1:], dtype=dt, device=dev)
                w["send"]["hdr"] = torch.zeros(max_episodes + 3, dtype=torch.int64, device=dev)  # E, N, offsets[max_episodes + 1]
                w["recv"]["hdr"] = torch.zeros((world, max_episodes + 3), dtype=torch.int64, device=dev)
                w["send"]["info"] = torch.zeros((max_episodes, cols_info), dtype=torch.int32, device=dev)
                w["recv"]["info"] = torch.zeros((world * max_episodes, cols_info), dtype=torch.int32, device=dev)
                ring["sets"].append(w)
            # constants of the offset arithmetic below
            ring["j"] = torch.arange(max_episodes, device=dev).reshape(1, -1)
            ring["rank_base"] = torch.arange(world, device=dev).reshape(-1, 1) * max_experiences
            ring["end"] = torch.full((1,), world * max_experiences, dtype=torch.int64, device=dev)
            cls.cache[key] = ring
        w = ring["sets"][ring["next"]]
        ring["next"] = (ring["next"] + 1) % cls.SLOTS
        return w, ring


def _all_gather_padded(batch, max_episodes, max_experiences, group, fields):
    """compact=False path: every field is gathered straight into its final [world * quota, ...] array (the layout bg_learner_update reads with
    explicit episode lengths): no host synchronisation, no unpacking copies, buffers reused across calls (two sets, alternating); on NCCL the
    per-field all-gathers are issued as ONE coalesced group (one launch, one host dispatch)."""
    from .episode import EpisodeBatch

    E, N = int(batch.n_episodes), int(batch.n_experiences)
    dev = batch.after_boards.device
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    cols_info = batch.ep_info.shape[1]
    w, ring = _GatherWorkspace.get(dev, world, max_episodes, max_experiences, cols_info)
    names = [n for n, _, _ in _EP_FIELDS if fields is None or n in fields]
    for name in names:
        if N:
            w["send"][name][:N].copy_(getattr(batch, name)[:N])
    hdr = w["send"]["hdr"]
    hdr[0], hdr[1] = E, N
    hdr[2:3 + E].copy_(batch.ep_offsets[: E + 1])
    if E:
        w["send"]["info"][:E].copy_(batch.ep_info[:E])
    pairs = [(w["recv"][n], w["send"][n]) for n in names] + [(w["recv"]["hdr"], hdr), (w["recv"]["info"], w["send"]["info"])]
    if world > 1:
        byte_pairs = [(out.reshape(-1).view(torch.uint8), inp.reshape(-1).view(torch.uint8)) for out, inp in pairs]  # int16 is not a collective dtype
        if dev.type == "cuda":
            done = False
            cm = getattr(dist, "_coalescing_manager", None)
            if cm is not None and _GatherWorkspace.coalescing:
                try:
                    with cm(group=group, device=dev, async_ops=False):
                        for ob, ib in byte_pairs:
                            dist.all_gather_into_tensor(ob, ib, group=group)
                    done = True
                except (RuntimeError, TypeError, NotImplementedError):  # a backend / version without coalesced all-gather: one call per field
                    _GatherWorkspace.coalescing = False
            if not done:
                for ob, ib in byte_pairs:
                    dist.all_gather_into_tensor(ob, ib, group=group)
        else:  # gloo (CPU tests): list form
            for ob, ib in byte_pairs:
                dist.all_gather(list(ob.reshape(world, -1).unbind(0)), ib, group=group)
    else:
        for out, inp in pairs:
            out.reshape(inp.shape).copy_(inp)
    H = w["recv"]["hdr"]
    offs = H[:, 2:]
    ep_len = torch.where(ring["j"] < H[:, 0:1], offs[:, 1:] - offs[:, :-1], 0).clamp_(min=0)
    ep_offsets = torch.cat([(offs[:, :-1] + ring["rank_base"]).reshape(-1), ring["end"]])
    g = lambda n: w["recv"][n] if n in names else None  # noqa: E731
    return EpisodeBatch(world * max_episodes, world * max_experiences, g("after_boards"), g("meta"), g("reward"), g("state_value"),
                        g("next_state_value"), g("n_moves"), g("action"), g("roll"), ep_offsets.contiguous(), w["recv"]["info"],
                        ep_len=ep_len.reshape(-1).to(torch.int32).contiguous())


LEARNER_FIELDS = ("after_boards", "meta", "reward")  # what Trainer.update / bg_learner_update(records=1) reads


def all_gather_episodes(batch, max_episodes: int, max_experiences: int, group=None, compact: bool = True, fields=None):
    """Config 5 (collective 3): every rank contributes its drained episodes (at most max_episodes /
    max_experiences) and receives all of them in rank order as one EpisodeBatch -- what the trainer rank feeds to Trainer.update.
    Compact records (72 B per experience, so 200 episodes are ~1.3 MB in total); replaces the reference's ExperienceQueue.put from every
    worker process (src/multi/worker.py:60-64).

    compact=True : ONE all_gather of a fixed-size byte buffer per rank; the result is a dense CSR batch (one small host read-back for the
                   per-rank sizes).
    compact=False: no host synchronisation and no unpacking copies -- each field is gathered straight into its final padded array (buffers
                   are cached, two sets in turn: the returned batch is valid until the SECOND next call with the same quota) and the batch carries explicit episode
                   lengths (EpisodeBatch.ep_len; bg_learner_update's ep_len argument); n_episodes = world * max_episodes, ranks that supplied
                   fewer episodes contribute zero-length ones, which the learner skips.  fields=LEARNER_FIELDS gathers only what the trainer
                   reads (the other record fields are None)."""
    from .episode import EpisodeBatch

    E, N = int(batch.n_episodes), int(batch.n_experiences)
    if E > max_episodes or N > max_experiences:
        raise ValueError(f"batch ({E} episodes, {N} experiences) exceeds the gather quota ({max_episodes}, {max_experiences})")
    if not compact:
        return _all_gather_padded(batch, max_episodes, max_experiences, group, fields)
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
