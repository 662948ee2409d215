"""Checkpoint-vs-checkpoint match runner on the GPU arena.

The reference only has an interactive front-end whose agent plays greedily (src/play/play_versus_ai.py:165-195: argmax of the
afterstate values); it has no way to measure playing strength.  This runs thousands of games between two weight sets at once,
with the reference agent's decision rule, using only kernels that already exist: bg_movegen over the arena's live positions,
bg_eval once per net, bg_select, and bg_arena_step(forced_action) to apply each side's choice (env.step(action))."""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .arena import Arena

POINTS = {1: 1, 2: 2, 3: 3}  # regular / gammon / backgammon in cube-less match points


def play_match(weights_a, weights_b, n_games: int = 4096, hidden_size: int = 128, device=None, seed: int = 0, temperature: float = 0.0,
               max_plies: int = 300, two_ply_a: Optional[tuple] = None, two_ply_b: Optional[tuple] = None, return_batch: bool = False) -> dict:
    """Play n_games games, A against B.  A sits as PLAYER1 in even games and PLAYER2 in odd games; dice are Philox (seed).
    weights_*: state_dict or packed tensor.  temperature 0 = the reference agent (greedy).  two_ply_* = (top_k, alpha, beta)
    makes that side rescore EVERY candidate with bg_two_ply (reference two_ply.py:44-150 semantics) before choosing.
    Note that the reference's value net is trained on un-signed rewards (the mover's win reward, no perspective flip in the TD
    target, trainer.py:110-115), so "more training" need not mean "wins more" under its own greedy rule; this runner measures, it
    does not assume.
    -> {"games", "a_wins", "b_wins", "unfinished", "a_win_rate", "a_points_per_game", "win_types_a", "win_types_b", "plies"}
    (+ "batch": the drained EpisodeBatch with per-decision mover/action traces, if return_batch)"""
    if not torch.cuda.is_available():
        raise RuntimeError("play_match needs a CUDA device (there is no CPU fallback)")
    dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    H = int(hidden_size)
    pk = [(w if isinstance(w, torch.Tensor) else ops.pack_weights(w)).to(dev, torch.float32).contiguous() for w in (weights_a, weights_b)]
    prep = [ops.prepare_weights(p, H) for p in pk]
    G = int(n_games)
    ar = Arena(G, hidden_size=H, device=dev, max_plies=max_plies, seed=seed, auto_reset=False, ring_experiences=G * max_plies,
               ring_episodes=max(G, 16))
    ar.set_weights(pk[0], version=1, temperature=temperature)  # unused by forced moves, needed by the arena's own evaluation pass
    ar.reset()
    a_is_p1 = (torch.arange(G, device=dev) % 2 == 0)
    pool = torch.empty((max(G * 40, 1 << 16), 52), dtype=torch.int8, device=dev)
    plies = 0
    while plies < max_plies:
        boards, players, rolls, gstate = ar.state()
        if plies % 16 == 0 and int((gstate == 0).sum().item()) == 0:
            break
        res = ops.movegen(boards, players, rolls, item_cap=ar.move_cap, out_boards=pool, want_owner=True, check_status=False)
        n_dev = res.total_dev
        side_a_item = (players == 0) == a_is_p1  # the mover of this game is A
        v = []
        for k in (0, 1):
            vk = ops.evaluate(pool, res.flags, prep[k], n_dev=n_dev)
            tp = (two_ply_a, two_ply_b)[k]
            if tp is not None:
                total = res.total
                rows = torch.nonzero(side_a_item[res.owner[:total].long()] == (k == 0)).squeeze(1)
                if rows.numel():
                    sc, _ = ops.two_ply(pool[rows].contiguous(), res.flags[rows].contiguous(), vk[rows].contiguous(), prep[k], top_k=tp[0], alpha=tp[1],
                                        beta=tp[2])
                    vk = vk.clone()
                    vk[rows] = sc
            v.append(vk)
        row_a = side_a_item[res.owner.long().clamp_(0, G - 1)]
        vals = torch.where(row_a, v[0], v[1])
        act = ops.select(vals, res.offsets, res.counts, temperature=temperature, seed=seed, ctr=plies, item_cap=ar.move_cap)
        ar.step(1, forced_action=act)
        plies += 1
    batch = ar.drain(max_episodes=G)
    ar.close()
    info = batch.ep_info[: batch.n_episodes].cpu()
    gid, winner, wtype = info[:, 9].long(), info[:, 1], info[:, 0]
    a_seat = torch.where(gid % 2 == 0, 0, 1).to(winner.dtype)
    decided = winner >= 0
    a_won = decided & (winner == a_seat)
    b_won = decided & (winner != a_seat)
    pts = torch.tensor([0, 1, 2, 3])[wtype.long().clamp(0, 3)]
    n = int(batch.n_episodes)
    out = {"games": n, "a_wins": int(a_won.sum()), "b_wins": int(b_won.sum()), "unfinished": G - int(decided.sum()), "plies": plies,
           "a_win_rate": float(a_won.sum()) / max(int(decided.sum()), 1),
           "a_points_per_game": float((pts * a_won).sum() - (pts * b_won).sum()) / max(n, 1),
           "win_types_a": {name: int((a_won & (wtype == k)).sum()) for k, name in ((1, "regular"), (2, "gammon"), (3, "backgammon"))},
           "win_types_b": {name: int((b_won & (wtype == k)).sum()) for k, name in ((1, "regular"), (2, "gammon"), (3, "backgammon"))}}
    if return_batch:
        out["batch"] = batch
    return out


def select_highest_value_action(policy_network, x: torch.Tensor) -> int:
    """The reference agent's decision rule (src/play/play_versus_ai.py:188-195): argmax of the afterstate values, lowest index on ties."""
    with torch.no_grad():
        return int(torch.argmax(policy_network.forward(x)).item())


def agent_play_step(policy_network, env) -> int:
    """src/play/play_versus_ai.py:165-185 without the printing: the action index the agent picks in `env`'s current position."""
    return select_highest_value_action(policy_network, env.legal_board_features[: env.num_moves])
