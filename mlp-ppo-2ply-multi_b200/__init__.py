"""B200-native (sm_100a) backgammon self-play hot path: legal-move generation, 198-feature afterstate encoding,
sigmoid-MLP value evaluation, softmax(V/T) action selection and 2-ply lookahead, behind the Python surface of
Nick-qsv/MLP-PPO-2PLY-MULTI.  All compute is in libbgarena.so (csrc/, hand-written CUDA); there is no CPU fallback."""
from . import _lib, arena, board, env, episode, match, moves, ops, parameter_manager, policy_network, trainer, types
from .arena import Arena, temperature_for_version
from .board import ImmutableBoard
from .env import BackgammonEnv
from .episode import Episode, EpisodeBatch, Experience
from .moves import execute_full_move_on_board_copy, generate_all_board_features, get_all_possible_moves
from .parameter_manager import ParameterManager
from .match import agent_play_step, play_match, select_highest_value_action
from .policy_network import BackgammonPolicyNetwork
from .trainer import TD0Learner, Trainer, features_to_boards
from .types import BoardState, FullMove, Player, Position, SubMove
from ._lib import BgError, SO_PATH
from .ops import (DICE_ROLLS, HostPipeline, MovegenResult, PreparedWeights, encode, evaluate, CompactResult, evaluate_codes, movegen, movegen_all_rolls, movegen_all_rolls_compact, movegen_evaluate,
                  movegen_evaluate_all_rolls,
                  pack_weights, prepare_weights, select, set_reply_sampling, two_ply, unpack_weights)

__all__ = ["HostPipeline", "agent_play_step", "select_highest_value_action", "play_match", "Trainer", "TD0Learner", "features_to_boards", "ImmutableBoard", "BackgammonEnv", "ParameterManager", "BackgammonPolicyNetwork", "execute_full_move_on_board_copy",
           "generate_all_board_features", "get_all_possible_moves", "Arena", "temperature_for_version", "Episode", "EpisodeBatch", "Experience", "BoardState", "FullMove", "Player",
           "Position", "SubMove", "ops", "BgError", "SO_PATH", "DICE_ROLLS", "MovegenResult", "PreparedWeights", "encode", "evaluate", "CompactResult", "evaluate_codes", "movegen_all_rolls_compact", "movegen", "movegen_all_rolls", "movegen_evaluate", "movegen_evaluate_all_rolls",
           "pack_weights", "prepare_weights", "select", "set_reply_sampling", "two_ply", "unpack_weights"]
