"""Legal-next-state API with the reference's names and result order, computed by the CUDA move generator:
get_all_possible_moves (reference src/backgammon/moves/generate_all_moves.py:7-66), generate_all_board_features and
execute_full_move_on_board_copy (src/environments/env_helper.py:7-91)."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch

from . import ops
from .board import ImmutableBoard, _device
from .types import FullMove, Player, Position, submoves_to_fullmove


def legal_moves_and_boards(player: Player, board: ImmutableBoard, roll_result) -> Tuple[List[FullMove], np.ndarray]:
    """(FullMove list in the reference's order, afterstate boards int8[n,52])"""
    dev = _device()
    b = torch.from_numpy(board.to_array()).to(dev).reshape(1, 52)
    p = torch.tensor([int(player)], dtype=torch.uint8, device=dev)
    r = torch.tensor([[int(roll_result[0]), int(roll_result[1])]], dtype=torch.uint8, device=dev)
    res = ops.movegen(b, p, r, item_cap=4096, pool_cap=4096, want_submoves=True, want_owner=False, want_flags=False)
    n = int(res.counts[0].item())
    off = int(res.offsets[0].item())
    sm = res.submoves[off:off + n].cpu().numpy()
    boards = res.boards[off:off + n].cpu().numpy()
    return [submoves_to_fullmove(m, player) for m in sm], boards


def get_all_possible_moves(player: Player, board: ImmutableBoard, roll_result) -> List[FullMove]:
    return legal_moves_and_boards(player, board, roll_result)[0]


def execute_full_move_on_board_copy(board: ImmutableBoard, full_move: FullMove) -> ImmutableBoard:
    """List-based application of a FullMove, no validation, hits_blot trusted (env_helper.py:27-91)."""
    me, opp = int(full_move.player), 1 - int(full_move.player)
    p = [list(board.positions_0), list(board.positions_1)]
    bar, off = list(board.bar), list(board.borne_off)
    for sm in full_move.sub_move_commands:
        s, e = int(sm.start), int(sm.end)
        if s == Position.BAR:
            bar[me] -= 1
        else:
            p[me][s] -= 1
        if sm.hits_blot:
            p[opp][e] -= 1
            bar[opp] += 1
        if e == Position.BEAR_OFF:
            off[me] += 1
        else:
            p[me][e] += 1
    return ImmutableBoard(tuple(p[0]), tuple(p[1]), tuple(bar), tuple(off), board.device)


def generate_all_board_features(board: ImmutableBoard, current_player: Player, legal_moves: List[FullMove]) -> torch.Tensor:
    """fp32 [N,198] features of every afterstate, flag = current_player (env_helper.py:7-24); encoded on the GPU."""
    if not legal_moves:
        return torch.zeros(0, 198, dtype=torch.float32, device=board.device)
    dev = _device()
    boards = np.stack([execute_full_move_on_board_copy(board, m).to_array() for m in legal_moves])
    flags = torch.full((len(legal_moves),), int(current_player), dtype=torch.uint8, device=dev)
    return ops.encode(torch.from_numpy(boards).to(dev), flags).to(board.device)
