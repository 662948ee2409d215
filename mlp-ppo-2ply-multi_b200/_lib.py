"""ctypes binding of libbgarena.so (include/bgarena.h).  There is no fallback: if the CUDA library is missing or a call
fails, this module raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("BG_LIBBGARENA") or os.path.join(_HERE, "libbgarena.so")  # the override is for A/B builds during development

BG_OK = 0
BG_ERR_ARG, BG_ERR_CUDA, BG_ERR_CAPACITY, BG_ERR_INVARIANT = -1, -2, -3, -4
BOARD_BYTES = 52
NUM_FEATURES = 198


class BgError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"libbgarena status {status}: {msg}")
        self.status = status


_lib = None

_vp, _i32, _i64, _u64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float

# name -> (restype, argtypes); mirrors include/bgarena.h one to one
SIGNATURES = {
    "bg_abi_version": (_i32, []),
    "bg_last_error": (C.c_char_p, []),
    "bg_device_count": (_i32, []),
    "bg_movegen_workspace_bytes": (_i64, [_i64]),
    "bg_movegen": (_i32, [_vp, _vp, _vp, _i64, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "bg_encode": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "bg_prepared_weights_bytes": (_i64, [_i32]),
    "bg_prepare_weights": (_i32, [_vp, _i32, _vp, _vp]),
    "bg_eval": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _i32, _vp, _vp]),
    "bg_eval_indirect": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _i32, _vp, _vp]),
    "bg_movegen_eval": (_i32, [_vp, _vp, _vp, _i64, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i32, _vp, _vp]),
    "bg_movegen_all_rolls": (_i32, [_vp, _vp, _i64, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "bg_movegen_eval_all_rolls": (_i32, [_vp, _vp, _i64, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i32, _vp, _vp]),
    "bg_movegen_all_rolls_compact": (_i32, [_vp, _vp, _i64, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "bg_eval_codes": (_i32, [_vp, _vp, _vp, _i64, _vp, _i64, _vp, _i32, _vp, _vp]),
    "bg_movegen_eval_all_rolls_compact": (_i32, [_vp, _vp, _i64, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i32, _vp, _vp]),
    "bg_afterstates_from_codes": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "bg_eval_tc_status": (_i32, []),
    "bg_eval_tc_tile_schedule": (_i32, [_i32]),
    "bg_select": (_i32, [_vp, _vp, _vp, _i32, _i64, _f32, _u64, _u64, _i64, _vp, _vp]),
    "bg_two_ply_workspace_bytes": (_i64, [_i64]),
    "bg_two_ply_reply_sampling": (_i32, [_i32, _u64]),
    "bg_two_ply": (_i32, [_vp, _vp, _vp, _i64, _vp, _i32, _i32, _f32, _f32, _vp, _vp, _vp, _vp, _i64, _vp]),
    "bg_hostpipe_create": (_i32, [_vp, _i32, _i32, _i64, _i32, _i32, _i32, _i32]),
    "bg_hostpipe_destroy": (_i32, [_vp]),
    "bg_hostpipe_run": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _f32, _u64, _vp, _vp, _vp]),
    "bg_hostpipe_status": (_i32, [_vp, _vp]),
    "bg_arena_create": (_i32, [_vp, _i32, _i64, _i32, _i32, _i32, _u64, _i64, _i64, _i64, _i32]),
    "bg_arena_destroy": (_i32, [_vp]),
    "bg_arena_set_weights": (_i32, [_vp, _vp, _i64, _f32, _vp]),
    "bg_arena_set_dice_tape": (_i32, [_vp, _vp, _i64, _vp]),
    "bg_arena_reset": (_i32, [_vp, _vp]),
    "bg_arena_set_lookahead": (_i32, [_vp, _i32, _i32, _f32, _f32]),
    "bg_arena_step": (_i32, [_vp, _i32, _i32, _vp, _vp]),
    "bg_arena_drain_episodes": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "bg_arena_stats": (_i32, [_vp, _vp, _vp]),
    "bg_arena_export_state": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "bg_learner_create": (_i32, [_vp, _i32, _i32, _f32, _f32, _f32]),
    "bg_learner_destroy": (_i32, [_vp]),
    "bg_learner_set_parameters": (_i32, [_vp, _vp, _i32, _vp]),
    "bg_learner_get_parameters": (_i32, [_vp, _vp, _vp]),
    "bg_learner_get_optimizer": (_i32, [_vp, _vp, _vp, _vp, _vp]),
    "bg_learner_update": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp]),
}


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(
                f"{SO_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback."
            )
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status: int):
    if status != BG_OK:
        raise BgError(status, (lib().bg_last_error() or b"").decode())
