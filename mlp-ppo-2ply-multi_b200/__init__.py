"""B200-native (sm_100a) backgammon self-play hot path: legal-move generation, 198-feature afterstate encoding,
sigmoid-MLP value evaluation, softmax(V/T) action selection and 2-ply lookahead, behind the Python surface of
Nick-qsv/MLP-PPO-2PLY-MULTI.  All compute is in libbgarena.so (csrc/, hand-written CUDA); there is no CPU fallback."""
from . import _lib, arena, episode, ops, types
from .arena import Arena, temperature_for_version
from .episode import Episode, EpisodeBatch, Experience
from .types import BoardState, FullMove, Player, Position, SubMove
from ._lib import BgError, SO_PATH
from .ops import (DICE_ROLLS, MovegenResult, PreparedWeights, encode, evaluate, movegen, pack_weights, prepare_weights, select,
                  two_ply, unpack_weights)

__all__ = ["Arena", "temperature_for_version", "Episode", "EpisodeBatch", "Experience", "BoardState", "FullMove", "Player",
           "Position", "SubMove", "ops", "BgError", "SO_PATH", "DICE_ROLLS", "MovegenResult", "PreparedWeights", "encode", "evaluate", "movegen",
           "pack_weights", "prepare_weights", "select", "two_ply", "unpack_weights"]
