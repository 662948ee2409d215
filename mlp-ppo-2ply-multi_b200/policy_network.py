"""BackgammonPolicyNetwork with the reference's state_dict and forward contract (reference src/agents/policy_network.py:36-70):
fc1 = Linear(198, H), sigmoid, value_head = Linear(H, 1), linear output, Xavier-uniform weights.  The module itself is plain
PyTorch (it is what the learner trains and what .pth checkpoints load into); `values()` evaluates boards through the fused
CUDA kernel instead of materialising features."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class BackgammonPolicyNetwork(nn.Module):
    def __init__(self, input_size: int = 198, hidden_size: int = 128):
        super().__init__()
        self.fc1 = nn.Linear(input_size, hidden_size)
        self.value_head = nn.Linear(hidden_size, 1)
        nn.init.xavier_uniform_(self.fc1.weight)
        nn.init.xavier_uniform_(self.value_head.weight)

    @property
    def hidden_size(self) -> int:
        return self.fc1.out_features

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [N,198] -> state values [N] (no output squashing)."""
        return self.value_head(torch.sigmoid(self.fc1(x))).squeeze(-1)

    # ---- CUDA fast path --------------------------------------------------------------------------------------------
    def packed(self, device=None) -> torch.Tensor:
        return ops.pack_weights(self.state_dict(), device=device)

    def prepared(self, device) -> ops.PreparedWeights:
        return ops.prepare_weights(self.packed(device), self.hidden_size)

    @torch.no_grad()
    def values(self, boards: torch.Tensor, flags: torch.Tensor, prepared: ops.PreparedWeights = None) -> torch.Tensor:
        """V(board, flag) for int8 [N,52] CUDA boards without building the 198-feature rows (bg_eval)."""
        w = prepared if prepared is not None else self.prepared(boards.device)
        return ops.evaluate(boards, flags, w)
