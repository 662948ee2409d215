"""ParameterManager with the reference's 4-method surface (reference src/multi/parameter_manager.py:54-111):
get_parameters / set_parameters / get_version / get_temperature, plus .pth save/load in the reference's checkpoint format.

What changes: instead of a multiprocessing.Manager dict polled by worker processes, set_parameters() publishes the packed
fp32 weight blob to every subscribed GPU arena -- directly on one GPU, and with ONE NCCL broadcast of the ~104 KB blob from the
trainer rank when torch.distributed is initialised; version and temperature are derived on every rank, so
publication never synchronises with the host."""
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

    def __init__(self, *reference_args, hidden_size: int = 128, src_rank: int = 0, process_group=None):
        """ParameterManager() or, as the reference constructs it (src/main.py:65-73), ParameterManager(lock, version, parameters): the
        three multiprocessing.Manager proxies are accepted and ignored (there are no worker processes to share them with).  A single int
        positional argument is taken as hidden_size."""
        if len(reference_args) == 1 and isinstance(reference_args[0], int):
            hidden_size = reference_args[0]
        elif len(reference_args) not in (0, 3):
            raise TypeError("ParameterManager() takes no positional arguments, or the reference's (lock, version, parameters)")
        self._lock = threading.Lock()
        self._version = 1
        self._hidden = hidden_size
        # the packed fp32 blob [W1^T | b1 | w2 | b2] is the source of truth (CPU or device tensor); state dicts are views of it
        self._packed = ops.pack_weights(BackgammonPolicyNetwork(hidden_size=hidden_size).state_dict()).detach().cpu()
        self._arenas: List = []
        self.src_rank = src_rank
        self.process_group = process_group
        self._bcast = None  # preallocated broadcast buffer (device)

    # ---- reference surface -----------------------------------------------------------------------------------------
    def get_parameters(self, device=None):
        sd = ops.unpack_weights(self._packed, self._hidden)
        return {k: v.to(device if device is not None else "cpu") for k, v in sd.items()}

    def get_version(self) -> int:
        return self._version

    def set_parameters(self, new_state_dict):
        self.set_packed(ops.pack_weights(new_state_dict), new_state_dict["fc1.weight"].shape[0])

    def set_packed(self, packed: torch.Tensor, hidden_size: Optional[int] = None):
        """set_parameters for an already packed (device) blob: what the CUDA learner hands over, no host round trip.
        COLLECTIVE when torch.distributed is initialised: every other rank calls sync_from_source() at the same point."""
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
        """LOCAL: the arena gets this rank's current weights now and every later publication.  (Ranks other than the source hold
        freshly initialised weights until the first set_packed() / sync_from_source() pair.)"""
        self._arenas.append(arena)
        arena.set_weights(self._packed.to(arena.device), version=self._version, temperature=self.get_temperature())

    def _distributed(self) -> bool:
        return torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size(self.process_group) > 1

    def _comm_device(self):
        if self._arenas:
            return self._arenas[0].device
        if self._packed.is_cuda:
            return self._packed.device
        backend = torch.distributed.get_backend(self.process_group)
        return torch.device("cuda", torch.cuda.current_device()) if "nccl" in str(backend) else torch.device("cpu")

    def publish(self):
        """Send the current weights to every subscribed arena; across ranks with ONE broadcast of the packed blob into a preallocated
        device buffer.  Version and temperature do not travel and nothing is read back on the host: the source bumps its version in
        set_packed(), every other rank bumps it in the matching sync_from_source(), and the temperature is a function of the version
        (parameter_manager.py:93-111 of the reference)."""
        if self._distributed():
            dev = self._comm_device()
            if self._bcast is None or self._bcast.numel() != self._packed.numel() or self._bcast.device != dev:
                self._bcast = torch.empty(self._packed.numel(), dtype=torch.float32, device=dev)
            if torch.distributed.get_rank(self.process_group) == self.src_rank:
                self._bcast.copy_(self._packed, non_blocking=True)
            torch.distributed.broadcast(self._bcast, src=self.src_rank, group=self.process_group)
            self._packed = self._bcast.clone()
        for a in self._arenas:
            a.set_weights(self._packed.to(a.device), version=self._version, temperature=self.get_temperature())

    def sync_from_source(self):
        """Ranks other than the source call this where the source rank calls set_packed() / set_parameters() (collective): receives the
        blob and advances the local version the same way."""
        if self._distributed() and torch.distributed.get_rank(self.process_group) != self.src_rank:
            with self._lock:
                self._version += 1
        self.publish()

    def check_version_sync(self) -> bool:
        """Debug aid (one host synchronisation): do all ranks agree on the version?"""
        if not self._distributed():
            return True
        dev = self._comm_device()
        v = torch.tensor([self._version, -self._version], dtype=torch.int64, device=dev)
        torch.distributed.all_reduce(v, op=torch.distributed.ReduceOp.MAX, group=self.process_group)
        lo_hi = v.tolist()
        return lo_hi[0] == -lo_hi[1]

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
