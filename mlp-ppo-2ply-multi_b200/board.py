"""ImmutableBoard value type with the reference's fields and methods (reference src/backgammon/board/immutable_board.py:16-258).
The 198-feature encoding runs on the GPU (bg_encode); there is no CPU encoder."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import numpy as np
import torch

from . import ops
from .types import Player, Position, SubMove, initial_board_array


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("this operation needs a CUDA device (libbgarena has no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


@dataclass(frozen=True)
class ImmutableBoard:
    positions_0: Tuple[int, ...]
    positions_1: Tuple[int, ...]
    bar: Tuple[int, int]
    borne_off: Tuple[int, int]
    device: torch.device = torch.device("cpu")  # device of the returned feature tensors (reference default)

    @staticmethod
    def initial_board(device: torch.device = torch.device("cpu")) -> "ImmutableBoard":
        return ImmutableBoard.from_array(initial_board_array(), device)

    @staticmethod
    def from_array(a, device: torch.device = torch.device("cpu")) -> "ImmutableBoard":
        a = [int(x) for x in a]
        return ImmutableBoard(tuple(a[0:24]), tuple(a[24:48]), (a[48], a[49]), (a[50], a[51]), device)

    def to_array(self) -> np.ndarray:
        return np.array(list(self.positions_0) + list(self.positions_1) + list(self.bar) + list(self.borne_off), np.int8)

    def __hash__(self):
        return hash((self.positions_0, self.positions_1, self.bar, self.borne_off))

    def get_board_features(self, current_player: Player) -> torch.Tensor:
        """fp32 [198] (immutable_board.py:86-128), computed by the CUDA encoder."""
        dev = _device()
        b = torch.from_numpy(self.to_array()).to(dev).reshape(1, 52)
        f = torch.tensor([int(current_player)], dtype=torch.uint8, device=dev)
        return ops.encode(b, f)[0].to(self.device)

    def move_checker(self, player: Player, sub_move: SubMove) -> "ImmutableBoard":
        """Apply one sub-move immutably; impossible input returns self (immutable_board.py:183-258)."""
        p = [list(self.positions_0), list(self.positions_1)]
        bar, off = list(self.bar), list(self.borne_off)
        me, opp = int(player), 1 - int(player)
        s, e = int(sub_move.start), int(sub_move.end)
        if s == Position.BAR:
            if bar[me] <= 0:
                return self
            bar[me] -= 1
        else:
            if p[me][s] <= 0:
                return self
            p[me][s] -= 1
        if sub_move.hits_blot:
            if p[opp][e] != 1:
                return self
            p[opp][e] -= 1
            bar[opp] += 1
        if e == Position.BEAR_OFF:
            off[me] += 1
        else:
            p[me][e] += 1
        return ImmutableBoard(tuple(p[0]), tuple(p[1]), tuple(bar), tuple(off), self.device)
