"""GPU-resident self-play arena: the replacement for the reference's worker processes
(reference src/multi/worker.py:47-179 + src/multi/experience_queue.py + ParameterManager polling).

    arena = Arena(n_games=65536, hidden_size=128, seed=0)
    arena.set_weights(net.state_dict(), version=1)      # temperature from the reference schedule unless given
    arena.reset()
    arena.step(32)                                       # 32 plies of every game, no host sync
    batch = arena.drain(max_episodes=200)                # EpisodeBatch (device) -> batch.to_episodes() for Trainer.update
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import ops
from ._lib import check, lib
from .episode import EpisodeBatch

# reference src/config/configuration.py:4,23-25
MAX_TIMESTEPS = 300
INITIAL_TEMPERATURE = 1.5
FINAL_TEMPERATURE = 0.5
MAX_UPDATES = 4000
MAX_LEGAL_MOVES = 500  # src/environments/backgammon_env.py:35

STAT_NAMES = ["games", "steps", "decisions", "passes", "afterstates", "win_regular", "win_gammon", "win_backgammon", "truncated",
              "p1_wins", "wait_steps", "errors", "replies"]
EP_INFO_INTS = 12


def temperature_for_version(version: int) -> float:
    """ParameterManager.get_temperature (reference src/multi/parameter_manager.py:93-111)."""
    if version <= 1:
        return INITIAL_TEMPERATURE
    if version >= 1 + MAX_UPDATES:
        return FINAL_TEMPERATURE
    return INITIAL_TEMPERATURE - (INITIAL_TEMPERATURE - FINAL_TEMPERATURE) * ((version - 1) / MAX_UPDATES)


class Arena:
    def __init__(self, n_games: int, hidden_size: int = 128, device=None, max_plies: int = MAX_TIMESTEPS, move_cap: int = MAX_LEGAL_MOVES,
                 seed: int = 0, game_id_base: int = 0, ring_experiences: int = 0, ring_episodes: int = 0, auto_reset: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("Arena needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.n_games, self.H, self.max_plies, self.move_cap = int(n_games), int(hidden_size), int(max_plies), int(move_cap)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().bg_arena_create(C.byref(self._h), self.device.index or 0, self.n_games, self.H, self.max_plies, self.move_cap,
                                        seed & (2**64 - 1), game_id_base, ring_experiences, ring_episodes, int(auto_reset)))
        self.version = 0
        self.temperature = INITIAL_TEMPERATURE
        self._packed = None

    # -- lifecycle -------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().bg_arena_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    # -- weights (ParameterManager.set_parameters side) ---------------------------------------------------------------
    def set_weights(self, weights, version: Optional[int] = None, temperature: Optional[float] = None):
        """weights: BackgammonPolicyNetwork.state_dict() or an already packed fp32 tensor (ops.pack_weights)."""
        packed = weights if isinstance(weights, torch.Tensor) else ops.pack_weights(weights)
        packed = packed.to(self.device, torch.float32).contiguous()
        if packed.numel() != 200 * self.H + 1:
            raise ValueError(f"weights do not match hidden_size={self.H}")
        self.version = self.version + 1 if version is None else int(version)
        self.temperature = temperature_for_version(self.version) if temperature is None else float(temperature)
        self._packed = packed  # keep alive until the prepare kernel has run
        check(lib().bg_arena_set_weights(self._h, packed.data_ptr(), self.version, self.temperature, self._stream()))

    def set_dice_tape(self, tape):
        """tape: uint8 [n_games, L, 2] (numpy or tensor), or None for Philox dice."""
        if tape is None:
            check(lib().bg_arena_set_dice_tape(self._h, None, 0, self._stream()))
            return
        t = torch.as_tensor(np.ascontiguousarray(tape) if isinstance(tape, np.ndarray) else tape, dtype=torch.uint8).contiguous()
        if t.dim() != 3 or t.shape[0] != self.n_games or t.shape[2] != 2:
            raise ValueError("tape must be [n_games, L, 2]")
        t = t.to(self.device)
        check(lib().bg_arena_set_dice_tape(self._h, t.data_ptr(), t.shape[1], self._stream()))
        torch.cuda.current_stream(self.device).synchronize()

    # -- play ---------------------------------------------------------------------------------------------------------
    def reset(self):
        check(lib().bg_arena_reset(self._h, self._stream()))

    def set_lookahead(self, n_candidates: int = 4, top_k: int = 5, alpha: float = 1.0, beta: float = 0.9):
        """2-ply policy parameters for step(lookahead=2).  Defaults = the reference's compute_scores_for_boards setting
        (src/multi/two_ply.py:44-90: top-4 candidates, mean of the top-5 replies, 1.0 * S - 0.9 * W); n_candidates=0 scores every
        legal afterstate (north_star expectimax, use top_k=1)."""
        check(lib().bg_arena_set_lookahead(self._h, int(n_candidates), int(top_k), float(alpha), float(beta)))

    def step(self, n_plies: int = 1, forced_action: Optional[torch.Tensor] = None, lookahead: int = 1):
        fa = None
        if forced_action is not None:
            fa = ops._req(forced_action, torch.int32, "forced_action")
            if fa.numel() != self.n_games:
                raise ValueError("forced_action must have n_games entries")
        check(lib().bg_arena_step(self._h, int(n_plies), int(lookahead), None if fa is None else fa.data_ptr(), self._stream()))

    def stats(self) -> dict:
        out = torch.zeros(16, dtype=torch.int64, device=self.device)
        check(lib().bg_arena_stats(self._h, out.data_ptr(), self._stream()))
        vals = out.cpu().tolist()
        return dict(zip(STAT_NAMES, vals))

    def state(self):
        """(boards int8[n,52], players uint8[n], rolls uint8[n,2], game_state uint8[n]) snapshot of the live games."""
        n, dev = self.n_games, self.device
        b = torch.empty((n, 52), dtype=torch.int8, device=dev)
        p = torch.empty(n, dtype=torch.uint8, device=dev)
        r = torch.empty((n, 2), dtype=torch.uint8, device=dev)
        s = torch.empty(n, dtype=torch.uint8, device=dev)
        check(lib().bg_arena_export_state(self._h, b.data_ptr(), p.data_ptr(), r.data_ptr(), s.data_ptr(), self._stream()))
        return b, p, r, s

    # -- hand-off (ExperienceQueue.get side) --------------------------------------------------------------------------------
    def drain(self, max_episodes: int = 200, max_experiences: Optional[int] = None) -> EpisodeBatch:
        if max_experiences is None:
            max_experiences = max_episodes * self.max_plies
        dev = self.device
        after = torch.empty((max_experiences, 52), dtype=torch.int8, device=dev)
        meta = torch.empty(max_experiences, dtype=torch.uint8, device=dev)
        reward = torch.empty(max_experiences, dtype=torch.float32, device=dev)
        v = torch.empty(max_experiences, dtype=torch.float32, device=dev)
        vn = torch.empty(max_experiences, dtype=torch.float32, device=dev)
        nm = torch.empty(max_experiences, dtype=torch.int16, device=dev)
        ac = torch.empty(max_experiences, dtype=torch.int16, device=dev)
        rl = torch.empty((max_experiences, 2), dtype=torch.uint8, device=dev)
        off = torch.zeros(max_episodes + 1, dtype=torch.int64, device=dev)
        info = torch.empty((max_episodes, EP_INFO_INTS), dtype=torch.int32, device=dev)
        out_n = torch.zeros(2, dtype=torch.int64, device=dev)
        check(lib().bg_arena_drain_episodes(self._h, max_episodes, max_experiences, after.data_ptr(), meta.data_ptr(), reward.data_ptr(),
                                            v.data_ptr(), vn.data_ptr(), nm.data_ptr(), ac.data_ptr(), rl.data_ptr(), off.data_ptr(),
                                            info.data_ptr(), out_n.data_ptr(), self._stream()))
        ne, nx = out_n.cpu().tolist()
        return EpisodeBatch(ne, nx, after, meta, reward, v, vn, nm, ac, rl, off, info)
