"""Import shim for the UNMODIFIED Python reference (test infrastructure; only usable where
/root/reference exists, i.e. in the build container -- never on the GPU box).

Recipe: SURVEY.md appendix C.  Both <ref> and <ref>/src go on sys.path (the reference imports its
own modules under two names), and four third-party modules it needs at import time but which are
not installed (gym, tensorboardX, boto3, botocore) are stubbed in sys.modules.
"""
from __future__ import annotations

import os
import sys
import types

REF = os.environ.get("BG_REFERENCE_PATH", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "src", "backgammon"))


def install():
    if not available():
        raise RuntimeError(f"reference not found at {REF}")
    for p in (os.path.join(REF, "src"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")

        class Env:
            metadata = {}

        class _Space:
            def __init__(self, *a, **k):
                self.args, self.kwargs = a, k

        spaces = types.ModuleType("gym.spaces")
        spaces.Box = _Space
        spaces.Discrete = _Space
        gym.Env = Env
        gym.spaces = spaces
        sys.modules["gym"] = gym
        sys.modules["gym.spaces"] = spaces
    if "tensorboardX" not in sys.modules:
        tbx = types.ModuleType("tensorboardX")

        class SummaryWriter:
            def __init__(self, *a, **k):
                pass

            def add_scalar(self, *a, **k):
                pass

            def add_scalars(self, *a, **k):
                pass

            def add_histogram(self, *a, **k):
                pass

            def flush(self):
                pass

            def close(self):
                pass

        rw = types.ModuleType("tensorboardX.record_writer")

        class RecordWriter:
            def __init__(self, *a, **k):
                pass

        class S3RecordWriter(RecordWriter):
            pass

        rw.RecordWriter = RecordWriter
        rw.S3RecordWriter = S3RecordWriter
        tbx.SummaryWriter = SummaryWriter
        tbx.record_writer = rw
        sys.modules["tensorboardX"] = tbx
        sys.modules["tensorboardX.record_writer"] = rw
    if "boto3" not in sys.modules:
        sys.modules["boto3"] = types.ModuleType("boto3")
    if "botocore" not in sys.modules:
        bc = types.ModuleType("botocore")
        ex = types.ModuleType("botocore.exceptions")

        class ClientError(Exception):
            pass

        ex.ClientError = ClientError
        bc.exceptions = ex
        sys.modules["botocore"] = bc
        sys.modules["botocore.exceptions"] = ex


def board_to_array(b):
    import numpy as np

    return np.array(list(b.positions_0) + list(b.positions_1) + list(b.bar) + list(b.borne_off), np.int8)


def array_to_board(a):
    from src.backgammon.board.immutable_board import ImmutableBoard

    a = [int(x) for x in a]
    return ImmutableBoard(positions_0=tuple(a[0:24]), positions_1=tuple(a[24:48]), bar=tuple(a[48:50]), borne_off=tuple(a[50:52]))


def array_to_board_env(a):
    """same, for the env-side copy of the module (imported as `backgammon...`, no `src.` prefix)"""
    from backgammon.board.immutable_board import ImmutableBoard

    a = [int(x) for x in a]
    return ImmutableBoard(positions_0=tuple(a[0:24]), positions_1=tuple(a[24:48]), bar=tuple(a[48:50]), borne_off=tuple(a[50:52]))


def moves_to_array(moves):
    """List[FullMove] -> uint8[n,4,3] padded with (255,255,0)"""
    import numpy as np

    out = np.zeros((len(moves), 4, 3), np.uint8)
    out[:, :, 0:2] = 255
    for i, m in enumerate(moves):
        for j, s in enumerate(m.sub_move_commands):
            out[i, j] = (int(s.start), int(s.end), int(bool(s.hits_blot)))
    return out
