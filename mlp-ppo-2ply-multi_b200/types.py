"""Data model mirroring the reference's (src/backgammon/types/moves.py:7-65, src/backgammon/board/immutable_board.py:16-70)."""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum, IntEnum, auto
from typing import Tuple

import numpy as np

NUMBER_OF_POINTS = 24

Position = IntEnum("Position", {**{f"P_{i}": i for i in range(24)}, "BAR": 24, "BEAR_OFF": 25})


class Player(IntEnum):
    PLAYER1 = 0
    PLAYER2 = 1


@dataclass(frozen=True)
class SubMove:
    start: Position
    end: Position
    hits_blot: bool


@dataclass
class FullMove:
    sub_move_commands: Tuple[SubMove, ...]
    player: Player


class BoardState(Enum):
    NORMAL = auto()
    ON_BAR = auto()
    BEAR_OFF = auto()
    GAME_OVER = auto()


def initial_board_array() -> np.ndarray:
    """int8[52] start position (immutable_board.py:26-70)."""
    b = np.zeros(52, np.int8)
    b[0], b[11], b[16], b[18] = 2, 5, 3, 5
    b[24 + 23], b[24 + 12], b[24 + 7], b[24 + 5] = 2, 5, 3, 5
    return b


def submoves_to_fullmove(sm: np.ndarray, player: int) -> FullMove:
    """uint8[4,3] (255-padded) -> FullMove"""
    subs = tuple(SubMove(Position(int(s)), Position(int(e)), bool(h)) for s, e, h in sm if s != 255)
    return FullMove(sub_move_commands=subs, player=Player(int(player)))
