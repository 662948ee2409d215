"""Episode / Experience records with the reference's field names and conversion semantics
(reference src/environments/episode.py:5-84), plus the device-side batch the arena drains.

Trainer.update (reference src/agents/trainer.py:81-153) reads experience.observation ([198] tensor on the trainer device),
experience.reward (0-dim tensor), episode.experiences, episode.win_type, episode.close_out_counts and
episode.prime_reward_counts; the remaining fields are carried for API parity."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

WIN_TYPES = {0: None, 1: "regular", 2: "gammon", 3: "backgammon"}


class Experience:
    __slots__ = ("observation", "state_value", "reward", "done", "next_observation", "next_state_value")

    def __init__(self, observation, state_value, reward, done, next_observation, next_state_value):
        self.observation = observation
        self.state_value = state_value
        self.reward = reward
        self.done = done
        self.next_observation = next_observation
        self.next_state_value = next_state_value

    def _fields(self):
        return self.__slots__

    def to_numpy(self):
        for name in self._fields():
            val = getattr(self, name)
            if isinstance(val, torch.Tensor):
                setattr(self, name, val.cpu().numpy())

    def to_tensor(self, device=None):
        # same dtype mapping as the reference (episode.py:30-46): ndarray -> tensor, float -> fp32, int (and bool, which is an
        # int subclass and is tested first there) -> int64
        for name in self._fields():
            val = getattr(self, name)
            if isinstance(val, np.ndarray):
                setattr(self, name, torch.from_numpy(val).to(device))
            elif isinstance(val, float):
                setattr(self, name, torch.tensor(val, dtype=torch.float32, device=device))
            elif isinstance(val, int):
                setattr(self, name, torch.tensor(val, dtype=torch.int64, device=device))
            elif isinstance(val, torch.Tensor) and device is not None:
                setattr(self, name, val.to(device))


class Episode:
    def __init__(self):
        self.experiences: List[Experience] = []
        self.win_type: Optional[str] = None
        self.close_out_counts = {}
        self.prime_reward_counts = {}

    def add_experience(self, experience, info):
        self.experiences.append(experience)
        if info.get("win_type"):
            self.win_type = info["win_type"]
        player = info.get("current_player", None)
        if player is not None:
            self.close_out_counts.setdefault(player, 0)
            self.prime_reward_counts.setdefault(player, 0)
            if info.get("close_out_reward", False):
                self.close_out_counts[player] += 1
            if info.get("prime_reward", False):
                self.prime_reward_counts[player] += 1

    def to_numpy(self):
        for e in self.experiences:
            e.to_numpy()

    def to_tensor(self, device=None):
        for e in self.experiences:
            e.to_tensor(device=device)


@dataclass
class EpisodeBatch:
    """Finished episodes as drained from the arena (device tensors, CSR over episodes)."""

    n_episodes: int
    n_experiences: int
    after_boards: torch.Tensor  # int8 [N,52] board after the mover's move
    meta: torch.Tensor  # uint8 [N] bit0 mover, bit1 next flag, bit2 done, bit3 close-out, bit4 prime
    reward: torch.Tensor  # fp32 [N]
    state_value: torch.Tensor  # fp32 [N] V(observation)
    next_state_value: torch.Tensor  # fp32 [N] V(chosen afterstate)
    n_moves: torch.Tensor  # int16 [N]
    action: torch.Tensor  # int16 [N]
    roll: torch.Tensor  # uint8 [N,2]
    ep_offsets: torch.Tensor  # int64 [E+1]
    ep_info: torch.Tensor  # int32 [E,12]
    ep_len: Optional[torch.Tensor] = None  # int32 [E]: set when the episodes are NOT contiguous (ep_offsets[e] is then only the first row of e)

    def episode_lengths(self) -> torch.Tensor:
        E = self.n_episodes
        return self.ep_len[:E].to(torch.int64) if self.ep_len is not None else self.ep_offsets[1:E + 1] - self.ep_offsets[:E]

    def observation_boards(self):
        """(obs_boards int8[N,52], obs_flags uint8[N]): the board each decision was made on and the player to move."""
        from .types import initial_board_array

        N = self.n_experiences
        dev = self.after_boards.device
        obs = torch.empty_like(self.after_boards[:N])
        if N:
            obs[1:] = self.after_boards[: N - 1]
            first = self.ep_offsets[: self.n_episodes]
            first = first[self.episode_lengths() > 0]  # padded batches carry zero-length filler episodes whose offset may be out of range
            obs[first] = torch.from_numpy(initial_board_array()).to(dev)
        return obs, (self.meta[:N] & 1)

    def observations(self):
        """fp32 [N,198] observation features (reference worker.py:149-158 `observation`), encoded on the GPU."""
        from . import ops

        obs, flags = self.observation_boards()
        return ops.encode(obs, flags.contiguous())

    def next_observations(self):
        from . import ops

        N = self.n_experiences
        return ops.encode(self.after_boards[:N].contiguous(), ((self.meta[:N] >> 1) & 1).contiguous())

    def to_episodes(self, with_next_observation: bool = True) -> List[Episode]:
        """Materialise reference-compatible Episode objects whose tensors are views into two [N,198] device tensors
        (what main.py:129-130 `episode.to_tensor(device)` produces, without the per-field H2D copies)."""
        from .types import Player

        N, E = self.n_experiences, self.n_episodes
        obs = self.observations().unbind(0)
        nxt = self.next_observations().unbind(0) if with_next_observation else [None] * N
        off = self.ep_offsets[: E + 1].tolist()
        lens = self.episode_lengths().tolist()
        info = self.ep_info[:E].tolist()
        # 0-dim tensor views, the dtypes Episode.to_tensor(device) produces in the reference (float -> fp32, bool -> int64)
        sv = self.state_value[:N].unbind(0)
        nsv = self.next_state_value[:N].unbind(0)
        rew = self.reward[:N].unbind(0)
        done = ((self.meta[:N] >> 2) & 1).to(torch.int64).unbind(0)
        out = []
        for k in range(E):
            ep = Episode()
            inf = info[k]
            ep.win_type = WIN_TYPES[inf[0]]
            for p in (0, 1):
                if (inf[8] >> p) & 1:
                    ep.close_out_counts[Player(p)] = inf[4 + p]
                    ep.prime_reward_counts[Player(p)] = inf[6 + p]
            lo, hi = off[k], off[k] + lens[k]
            ep.experiences = [Experience(obs[t], sv[t], rew[t], done[t], nxt[t], nsv[t]) for t in range(lo, hi)]
            out.append(ep)
        return out
