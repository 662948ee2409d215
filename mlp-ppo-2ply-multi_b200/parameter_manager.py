"""ParameterManager with the reference's 4-method surface (reference src/multi/parameter_manager.py:54-111):
get_parameters / set_parameters / get_version / get_temperature, plus .pth save/load in the reference's checkpoint format.

What changes: instead of a multiprocessing.Manager dict polled by worker processes, set_parameters() publishes the packed
fp32 weight blob to every subscribed GPU arena -- directly on one GPU, and with ONE NCCL broadcast of the ~104 KB blob
(+ version, temperature) from the trainer rank when torch.distributed is initialised."""
from __future__ import annotations

import os
import threading
from typing import List, Optional

import torch

from . import ops
from .arena import FINAL_TEMPERATURE, INITIAL_TEMPERATURE, MAX_UPDATES, temperature_for_version
from .policy_network import BackgammonPolicyNetwork


class ParameterManager:
    INITIAL_TEMPERATURE = INITIAL_TEMPERATURE
    FINAL_TEMPERATURE = FINAL_TEMPERATURE
    MAX_UPDATES = MAX_UPDATES

    def __init__(self, hidden_size: int = 128, src_rank: int = 0, process_group=None):
        self._lock = threading.Lock()
        self._version = 1
        self._hidden = hidden_size
        # the packed fp32 blob [W1^T | b1 | w2 | b2] is the source of truth (CPU or device tensor); state dicts are views of it
        self._packed = ops.pack_weights(BackgammonPolicyNetwork(hidden_size=hidden_size).state_dict()).detach().cpu()
        self._arenas: List = []
        self.src_rank = src_rank
        self.process_group = process_group

    # ---- reference surface -----------------------------------------------------------------------------------------
    def get_parameters(self, device=None):
        sd = ops.unpack_weights(self._packed, self._hidden)
        return {k: v.to(device if device is not None else "cpu") for k, v in sd.items()}

    def get_version(self) -> int:
        return self._version

    def set_parameters(self, new_state_dict):
        self.set_packed(ops.pack_weights(new_state_dict), new_state_dict["fc1.weight"].shape[0])

    def set_packed(self, packed: torch.Tensor, hidden_size: Optional[int] = None):
        """set_parameters for an already packed (device) blob: what the CUDA learner hands over, no host round trip."""
        with self._lock:
            if hidden_size is not None:
                self._hidden = int(hidden_size)
            self._packed = packed.detach().to(torch.float32).reshape(-1).clone()
            self._version += 1
        self.publish()

    def get_temperature(self) -> float:
        return temperature_for_version(self.get_version())

    # ---- arena subscription / multi-GPU publication -------------------------------------------------------------------
    def subscribe(self, arena):
        self._arenas.append(arena)
        self.publish()

    def _distributed(self) -> bool:
        return torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size(self.process_group) > 1

    def publish(self):
        """Send the current weights (+version, temperature) to every subscribed arena; across ranks with one broadcast."""
        if not self._arenas and not self._distributed():
            return
        if self._distributed():
            dev = self._arenas[0].device if self._arenas else self._packed.device
            blob = torch.cat([self._packed.to(dev), torch.tensor([float(self._version), self.get_temperature()], device=dev)])
            torch.distributed.broadcast(blob, src=self.src_rank, group=self.process_group)
            self._version = int(blob[-2].item())
            self._packed = blob[:-2].clone()
        for a in self._arenas:
            a.set_weights(self._packed.to(a.device), version=self._version, temperature=self.get_temperature())

    def sync_from_source(self):
        """Non-source ranks call this where the source rank calls set_parameters()/publish() (collective)."""
        self.publish()

    # ---- checkpoints (reference :115-230, S3 dropped) ------------------------------------------------------------------
    def save_model(self, filename: Optional[str] = None, to_s3: bool = False):
        if to_s3:
            raise NotImplementedError("S3 checkpoint IO is out of scope (SURVEY.md section 2, row 12); use a local path")
        os.makedirs("models", exist_ok=True)
        path = os.path.join("models", filename or "ppo_backgammon.pth")
        torch.save(self.get_parameters(), path)
        return path

    def load_model(self, filename: Optional[str] = None, from_s3: bool = False):
        if from_s3:
            raise NotImplementedError("S3 checkpoint IO is out of scope; use a local path")
        path = filename if filename and os.path.isabs(filename) else os.path.join("models", filename or "ppo_backgammon.pth")
        self.set_parameters(torch.load(path, map_location="cpu"))
