"""Tensor-level operators over the C ABI (include/bgarena.h).  PyTorch is used only for device memory and streams; all
compute happens in libbgarena.so's sm_100a kernels.  Every function raises if the library or a CUDA device is missing."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import BOARD_BYTES, NUM_FEATURES, check, lib

DICE_ROLLS = [(a, b) for a in range(1, 7) for b in range(a, 7)]  # reference src/multi/two_ply.py:10-32
ROLL_COUNTS = [1 if a == b else 2 for a, b in DICE_ROLLS]  # src/multi/two_ply.py:33


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _req(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (libbgarena has no CPU path)")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    return t.contiguous()


def pack_weights(state_dict, device=None) -> torch.Tensor:
    """BackgammonPolicyNetwork.state_dict() (reference src/agents/policy_network.py:36-51: fc1.weight[H,198], fc1.bias[H],
    value_head.weight[1,H], value_head.bias[1]) -> packed fp32 [W1^T (198,H) | b1 | w2 | b2]."""
    W1 = state_dict["fc1.weight"].detach().to(torch.float32)
    parts = [W1.t().contiguous().reshape(-1), state_dict["fc1.bias"].detach().reshape(-1).to(torch.float32),
             state_dict["value_head.weight"].detach().reshape(-1).to(torch.float32),
             state_dict["value_head.bias"].detach().reshape(-1).to(torch.float32)]
    if W1.shape[1] != NUM_FEATURES:
        raise ValueError(f"fc1.weight must be [H,{NUM_FEATURES}]")
    packed = torch.cat([p.to(device or W1.device) for p in parts])
    return packed


def unpack_weights(packed: torch.Tensor, H: int) -> dict:
    packed = packed.detach()
    o = NUM_FEATURES * H
    return {"fc1.weight": packed[:o].reshape(NUM_FEATURES, H).t().contiguous(), "fc1.bias": packed[o:o + H].clone(),
            "value_head.weight": packed[o + H:o + 2 * H].reshape(1, H).clone(), "value_head.bias": packed[o + 2 * H:o + 2 * H + 1].clone()}


@dataclass
class PreparedWeights:
    """Device-side weight table for the fused evaluator (bg_prepare_weights)."""

    table: torch.Tensor
    H: int


def prepare_weights(packed: torch.Tensor, H: Optional[int] = None, out: Optional[torch.Tensor] = None) -> PreparedWeights:
    packed = _req(packed, torch.float32, "packed")
    if H is None:
        H = (packed.numel() - 1) // 200
    if packed.numel() != 200 * H + 1:
        raise ValueError(f"packed weights must have 200*H+1 floats (H={H}), got {packed.numel()}")
    nbytes = lib().bg_prepared_weights_bytes(H)
    if out is None:
        out = torch.empty(nbytes // 4, dtype=torch.float32, device=packed.device)
    check(lib().bg_prepare_weights(packed.data_ptr(), H, out.data_ptr(), _stream()))
    return PreparedWeights(out, H)


@dataclass
class MovegenResult:
    boards: torch.Tensor  # int8 [pool_cap, 52]; valid rows: union of item segments
    submoves: Optional[torch.Tensor]  # uint8 [pool_cap, 4, 3] or None
    owner: Optional[torch.Tensor]  # int32 [pool_cap] item index of each row
    flags: Optional[torch.Tensor]  # uint8 [pool_cap] the item's player per row (feature flag of the afterstate)
    offsets: torch.Tensor  # int64 [B] segment start per item (-1 if the item failed)
    counts: torch.Tensor  # int32 [B] TRUE number of legal moves (may exceed item_cap)
    total_dev: torch.Tensor  # int64 [1] rows used
    status_dev: torch.Tensor  # int32 [1]
    item_cap: int

    @property
    def total(self) -> int:
        return int(self.total_dev.item())

    def raise_for_status(self):
        st = int(self.status_dev.item())
        if st != 0:
            raise _lib.BgError(st, "bg_movegen reported a capacity/invariant problem (pool_cap too small, >4096 moves, or an invalid board)")

    def canonical(self):
        """(offsets int64[B+1], boards[T,52], submoves[T,4,3]|None) in item order (CSR), item_cap applied."""
        kept = torch.clamp(self.counts.to(torch.int64), max=self.item_cap)
        kept = torch.where(self.offsets >= 0, kept, torch.zeros_like(kept))
        off = torch.zeros(kept.numel() + 1, dtype=torch.int64, device=kept.device)
        off[1:] = torch.cumsum(kept, 0)
        T = int(off[-1].item())
        item = torch.repeat_interleave(torch.arange(kept.numel(), device=kept.device), kept, output_size=T)
        rows = self.offsets[item] + (torch.arange(T, device=kept.device) - off[item])
        return off, self.boards[rows], (None if self.submoves is None else self.submoves[rows])


_workspaces: dict = {}


def _workspace(B: int, device) -> torch.Tensor:
    need = lib().bg_movegen_workspace_bytes(B)
    # one scratch buffer per (device, stream): two streams generating moves at the same time must not share overflow lists / counters
    key = (device.index if device.index is not None else torch.cuda.current_device(), torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def movegen(boards: torch.Tensor, players: torch.Tensor, rolls: torch.Tensor, item_cap: int = 500, pool_cap: Optional[int] = None,
            want_submoves: bool = False, want_owner: bool = True, check_status: bool = True, out_boards: Optional[torch.Tensor] = None,
            workspace: Optional[torch.Tensor] = None, out_owner: Optional[torch.Tensor] = None, want_flags: bool = True,
            out_flags: Optional[torch.Tensor] = None) -> MovegenResult:
    """Batched get_all_possible_moves + execute_full_move_on_board_copy (reference generate_all_moves.py:7, env_helper.py:27)."""
    boards = _req(boards, torch.int8, "boards").reshape(-1, BOARD_BYTES)
    B = boards.shape[0]
    players = _req(players, torch.uint8, "players").reshape(B)
    rolls = _req(rolls, torch.uint8, "rolls").reshape(B, 2)
    dev = boards.device
    if pool_cap is None:
        pool_cap = out_boards.shape[0] if out_boards is not None else max(1024, B * 64)
    if out_boards is None:
        out_boards = torch.empty((pool_cap, BOARD_BYTES), dtype=torch.int8, device=dev)
    sub = torch.empty((pool_cap, 4, 3), dtype=torch.uint8, device=dev) if want_submoves else None
    owner = out_owner if out_owner is not None else (torch.empty(pool_cap, dtype=torch.int32, device=dev) if want_owner else None)
    flags = out_flags if out_flags is not None else (torch.empty(pool_cap, dtype=torch.uint8, device=dev) if want_flags else None)
    offsets = torch.empty(B, dtype=torch.int64, device=dev)
    counts = torch.empty(B, dtype=torch.int32, device=dev)
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = workspace if workspace is not None else _workspace(B, dev)
    check(lib().bg_movegen(boards.data_ptr(), players.data_ptr(), rolls.data_ptr(), B, item_cap, pool_cap, out_boards.data_ptr(),
                           _ptr(sub), _ptr(owner), _ptr(flags), offsets.data_ptr(), counts.data_ptr(), total.data_ptr(), status.data_ptr(),
                           ws.data_ptr(), ws.numel(), _stream()))
    res = MovegenResult(out_boards, sub, owner, flags, offsets, counts, total, status, item_cap)
    if check_status:
        res.raise_for_status()
    return res


def encode(boards: torch.Tensor, flags: torch.Tensor) -> torch.Tensor:
    """[N,52] int8 boards -> [N,198] fp32 features (reference immutable_board.py:86-128), bit-exact."""
    boards = _req(boards, torch.int8, "boards").reshape(-1, BOARD_BYTES)
    N = boards.shape[0]
    flags = _req(flags, torch.uint8, "flags").reshape(N)
    out = torch.empty((N, NUM_FEATURES), dtype=torch.float32, device=boards.device)
    check(lib().bg_encode(boards.data_ptr(), flags.data_ptr(), N, out.data_ptr(), _stream()))
    return out


def evaluate(boards: torch.Tensor, flags: Optional[torch.Tensor], weights: PreparedWeights, owner: Optional[torch.Tensor] = None,
             owner_players: Optional[torch.Tensor] = None, n_dev: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fused encode + BackgammonPolicyNetwork.forward (reference policy_network.py:53-70) straight from boards."""
    boards = _req(boards, torch.int8, "boards").reshape(-1, BOARD_BYTES)
    N = boards.shape[0]
    if flags is not None:
        flags = _req(flags, torch.uint8, "flags").reshape(N)
    else:
        owner = _req(owner, torch.int32, "owner")
        owner_players = _req(owner_players, torch.uint8, "owner_players")
    if out is None:
        out = torch.empty(N, dtype=torch.float32, device=boards.device)
    if n_dev is None:
        check(lib().bg_eval(boards.data_ptr(), _ptr(flags), _ptr(owner), _ptr(owner_players), N, weights.table.data_ptr(), weights.H,
                            out.data_ptr(), _stream()))
    else:
        n_dev = _req(n_dev, torch.int64, "n_dev")
        check(lib().bg_eval_indirect(boards.data_ptr(), _ptr(flags), _ptr(owner), _ptr(owner_players), n_dev.data_ptr(), N,
                                     weights.table.data_ptr(), weights.H, out.data_ptr(), _stream()))
    return out


def movegen_evaluate(boards: torch.Tensor, players: torch.Tensor, rolls: torch.Tensor, weights: PreparedWeights, out_boards: torch.Tensor,
                     out_flags: torch.Tensor, out_values: torch.Tensor, workspace: Optional[torch.Tensor] = None, item_cap: int = 500,
                     check_status: bool = False):
    """bg_movegen + bg_eval over the whole pool in one call (bg_movegen_eval): the evaluation of the bulk tier's afterstates overlaps the
    generation of the tail tiers.  -> (MovegenResult, values) with values[row] valid for the pool rows of every item."""
    boards = _req(boards, torch.int8, "boards").reshape(-1, BOARD_BYTES)
    B = boards.shape[0]
    players = _req(players, torch.uint8, "players").reshape(B)
    rolls = _req(rolls, torch.uint8, "rolls").reshape(B, 2)
    dev = boards.device
    pool_cap = out_boards.shape[0]
    if out_flags.numel() < pool_cap or out_values.numel() < pool_cap:
        raise ValueError("out_flags / out_values must have pool_cap entries")
    offsets = torch.empty(B, dtype=torch.int64, device=dev)
    counts = torch.empty(B, dtype=torch.int32, device=dev)
    total = torch.zeros(2, dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = workspace if workspace is not None else _workspace(B, dev)
    check(lib().bg_movegen_eval(boards.data_ptr(), players.data_ptr(), rolls.data_ptr(), B, item_cap, pool_cap, out_boards.data_ptr(),
                                out_flags.data_ptr(), offsets.data_ptr(), counts.data_ptr(), total.data_ptr(), status.data_ptr(), ws.data_ptr(),
                                ws.numel(), weights.table.data_ptr(), weights.H, out_values.data_ptr(), _stream()))
    res = MovegenResult(out_boards, None, None, out_flags, offsets, counts, total[:1], status, item_cap)
    if check_status:
        res.raise_for_status()
    return res, out_values


def movegen_all_rolls(boards: torch.Tensor, players: torch.Tensor, item_cap: int = 500, pool_cap: Optional[int] = None,
                      want_submoves: bool = False, want_owner: bool = False, check_status: bool = True, out_boards: Optional[torch.Tensor] = None,
                      workspace: Optional[torch.Tensor] = None, want_flags: bool = True, out_flags: Optional[torch.Tensor] = None) -> MovegenResult:
    """get_all_possible_moves for EVERY roll of DICE_ROLLS (reference src/multi/two_ply.py:10-32) of each position: item
    p * 21 + r.  One warp expands a whole position (bg_movegen_all_rolls); same results as movegen() on the replicated items."""
    boards = _req(boards, torch.int8, "boards").reshape(-1, BOARD_BYTES)
    P = boards.shape[0]
    B = 21 * P
    players = _req(players, torch.uint8, "players").reshape(P)
    dev = boards.device
    if pool_cap is None:
        pool_cap = out_boards.shape[0] if out_boards is not None else max(1024, B * 64)
    if out_boards is None:
        out_boards = torch.empty((pool_cap, BOARD_BYTES), dtype=torch.int8, device=dev)
    sub = torch.empty((pool_cap, 4, 3), dtype=torch.uint8, device=dev) if want_submoves else None
    owner = torch.empty(pool_cap, dtype=torch.int32, device=dev) if want_owner else None
    flags = out_flags if out_flags is not None else (torch.empty(pool_cap, dtype=torch.uint8, device=dev) if want_flags else None)
    offsets = torch.empty(B, dtype=torch.int64, device=dev)
    counts = torch.empty(B, dtype=torch.int32, device=dev)
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = workspace if workspace is not None else _workspace(B, dev)
    check(lib().bg_movegen_all_rolls(boards.data_ptr(), players.data_ptr(), P, item_cap, pool_cap, out_boards.data_ptr(), _ptr(sub), _ptr(owner),
                                     _ptr(flags), offsets.data_ptr(), counts.data_ptr(), total.data_ptr(), status.data_ptr(), ws.data_ptr(),
                                     ws.numel(), _stream()))
    res = MovegenResult(out_boards, sub, owner, flags, offsets, counts, total, status, item_cap)
    if check_status:
        res.raise_for_status()
    return res


def movegen_evaluate_all_rolls(boards: torch.Tensor, players: torch.Tensor, weights: PreparedWeights, out_boards: torch.Tensor,
                               out_flags: torch.Tensor, out_values: torch.Tensor, workspace: Optional[torch.Tensor] = None, item_cap: int = 500,
                               check_status: bool = False):
    """bg_movegen_eval_all_rolls: every roll of every position generated (one warp per position) and every afterstate evaluated."""
    boards = _req(boards, torch.int8, "boards").reshape(-1, BOARD_BYTES)
    P = boards.shape[0]
    B = 21 * P
    players = _req(players, torch.uint8, "players").reshape(P)
    dev = boards.device
    pool_cap = out_boards.shape[0]
    if out_flags.numel() < pool_cap or out_values.numel() < pool_cap:
        raise ValueError("out_flags / out_values must have pool_cap entries")
    offsets = torch.empty(B, dtype=torch.int64, device=dev)
    counts = torch.empty(B, dtype=torch.int32, device=dev)
    total = torch.zeros(2, dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = workspace if workspace is not None else _workspace(B, dev)
    check(lib().bg_movegen_eval_all_rolls(boards.data_ptr(), players.data_ptr(), P, item_cap, pool_cap, out_boards.data_ptr(),
                                          out_flags.data_ptr(), offsets.data_ptr(), counts.data_ptr(), total.data_ptr(), status.data_ptr(),
                                          ws.data_ptr(), ws.numel(), weights.table.data_ptr(), weights.H, out_values.data_ptr(), _stream()))
    res = MovegenResult(out_boards, None, None, out_flags, offsets, counts, total[:1], status, item_cap)
    if check_status:
        res.raise_for_status()
    return res, out_values


@dataclass
class CompactResult:
    """bg_movegen[_eval]_all_rolls_compact: one uint64 (code | position << 32) per legal afterstate instead of a board."""

    codes: torch.Tensor  # int64 [pool_cap] (bit pattern of the uint64 entries)
    offsets: torch.Tensor  # int64 [21 * P]
    counts: torch.Tensor  # int32 [21 * P] TRUE number of legal moves
    total_dev: torch.Tensor  # int64 [1]
    status_dev: torch.Tensor  # int32 [1]
    item_cap: int
    boards: torch.Tensor  # the POSITIONS' boards [P, 52] and players [P] the codes refer to
    players: torch.Tensor

    @property
    def total(self) -> int:
        return int(self.total_dev.item())

    def raise_for_status(self):
        st = int(self.status_dev.item())
        if st != 0:
            raise _lib.BgError(st, "bg_movegen_all_rolls_compact reported a capacity/invariant problem")

    def afterstates(self, rows: Optional[torch.Tensor] = None, n: Optional[int] = None) -> torch.Tensor:
        """boards [n,52] of the given pool rows (int64; negative -> zeros), or of rows 0..n-1"""
        if rows is not None:
            rows = _req(rows, torch.int64, "rows")
            n = rows.numel()
        out = torch.empty((n, BOARD_BYTES), dtype=torch.int8, device=self.codes.device)
        check(lib().bg_afterstates_from_codes(self.boards.data_ptr(), self.players.data_ptr(), self.codes.data_ptr(), _ptr(rows), n, out.data_ptr(), _stream()))
        return out

    def chosen_afterstates(self, actions: torch.Tensor) -> torch.Tensor:
        """the afterstate of each item's action (what env.step(action) moves to); zeros where action < 0"""
        a = actions.to(torch.int64)
        return self.afterstates(torch.where(a >= 0, self.offsets + a, torch.full_like(a, -1)))


def movegen_all_rolls_compact(boards: torch.Tensor, players: torch.Tensor, weights: Optional[PreparedWeights] = None, item_cap: int = 500,
                              pool_cap: Optional[int] = None, out_codes: Optional[torch.Tensor] = None, out_values: Optional[torch.Tensor] = None,
                              workspace: Optional[torch.Tensor] = None, check_status: bool = False):
    """Position-major move generation into a COMPACT pool (8 bytes per afterstate, no boards); with `weights` also the value of every
    afterstate (bg_movegen_eval_all_rolls_compact: the evaluator rebuilds each afterstate on chip).  -> CompactResult[, values]"""
    boards = _req(boards, torch.int8, "boards").reshape(-1, BOARD_BYTES)
    P = boards.shape[0]
    B = 21 * P
    players = _req(players, torch.uint8, "players").reshape(P)
    dev = boards.device
    if pool_cap is None:
        pool_cap = out_codes.numel() if out_codes is not None else max(1024, B * 64)
    if out_codes is None:
        out_codes = torch.empty(pool_cap, dtype=torch.int64, device=dev)
    offsets = torch.empty(B, dtype=torch.int64, device=dev)
    counts = torch.empty(B, dtype=torch.int32, device=dev)
    total = torch.zeros(2, dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = workspace if workspace is not None else _workspace(B, dev)
    if weights is None:
        check(lib().bg_movegen_all_rolls_compact(boards.data_ptr(), players.data_ptr(), P, item_cap, pool_cap, out_codes.data_ptr(), offsets.data_ptr(),
                                                 counts.data_ptr(), total.data_ptr(), status.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
    else:
        if out_values is None:
            out_values = torch.empty(pool_cap, dtype=torch.float32, device=dev)
        check(lib().bg_movegen_eval_all_rolls_compact(boards.data_ptr(), players.data_ptr(), P, item_cap, pool_cap, out_codes.data_ptr(),
                                                      offsets.data_ptr(), counts.data_ptr(), total.data_ptr(), status.data_ptr(), ws.data_ptr(),
                                                      ws.numel(), weights.table.data_ptr(), weights.H, out_values.data_ptr(), _stream()))
    res = CompactResult(out_codes, offsets, counts, total[:1], status, item_cap, boards, players)
    if check_status:
        res.raise_for_status()
    return res if weights is None else (res, out_values)


def evaluate_codes(res: CompactResult, weights: PreparedWeights, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """values of every row of a compact pool (bg_eval_codes, row count read on the device)"""
    if out is None:
        out = torch.empty(res.codes.numel(), dtype=torch.float32, device=res.codes.device)
    check(lib().bg_eval_codes(res.boards.data_ptr(), res.players.data_ptr(), res.codes.data_ptr(), 0, res.total_dev.data_ptr(), res.codes.numel(),
                              weights.table.data_ptr(), weights.H, out.data_ptr(), _stream()))
    return out


def select(values: torch.Tensor, offsets: torch.Tensor, counts: torch.Tensor, temperature: float, seed: int = 0, ctr: int = 0,
           item_cap: int = 500, item_id_base: int = 0) -> torch.Tensor:
    """softmax(V/T) sampling (reference worker.py:136-143) or, temperature <= 0, first-index argmax (play_versus_ai.py:188-195)."""
    values = _req(values, torch.float32, "values")
    offsets = _req(offsets, torch.int64, "offsets")
    counts = _req(counts, torch.int32, "counts")
    B = offsets.numel()
    out = torch.empty(B, dtype=torch.int32, device=values.device)
    check(lib().bg_select(values.data_ptr(), offsets.data_ptr(), counts.data_ptr(), item_cap, B, float(temperature), seed & (2**64 - 1),
                          ctr & (2**64 - 1), item_id_base, out.data_ptr(), _stream()))
    return out


def two_ply(cand_boards: torch.Tensor, mover: torch.Tensor, S: torch.Tensor, weights: PreparedWeights, top_k: int = 5, alpha: float = 1.0,
            beta: float = 0.9, workspace: Optional[torch.Tensor] = None, check_status: bool = True):
    """2-ply scores (reference src/multi/two_ply.py:44-150): alpha*S - beta*sum_r p_r*mean(top_k opponent replies).
    -> (scores fp32[N], replies int64[N])"""
    cand_boards = _req(cand_boards, torch.int8, "cand_boards").reshape(-1, BOARD_BYTES)
    N = cand_boards.shape[0]
    mover = _req(mover, torch.uint8, "mover").reshape(N)
    S = _req(S, torch.float32, "S").reshape(N)
    dev = cand_boards.device
    out = torch.empty(N, dtype=torch.float32, device=dev)
    nrep = torch.zeros(N, dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    if workspace is None:
        workspace = torch.empty(lib().bg_two_ply_workspace_bytes(N), dtype=torch.uint8, device=dev)
    check(lib().bg_two_ply(cand_boards.data_ptr(), mover.data_ptr(), S.data_ptr(), N, weights.table.data_ptr(), weights.H, top_k, alpha, beta,
                           out.data_ptr(), nrep.data_ptr(), status.data_ptr(), workspace.data_ptr(), workspace.numel(), _stream()))
    if check_status:
        st = int(status.item())
        if st != 0:
            raise _lib.BgError(st, "bg_two_ply: reply pool capacity exceeded for some candidates")
    return out, nrep


def set_reply_sampling(cap: int = 0, seed: int = 0) -> int:
    """Process-wide option of the 2-ply scorer (two_ply, Arena.step(lookahead=2)): cap > 0 reproduces the reference's
    `random.sample(opponent_moves, 50)` on the rolls 1-1 / 2-2 / 3-3 (two_ply.py:119-121) as a reproducible uniform sample of `cap`
    replies taken BEFORE evaluation; 0 (default) evaluates every reply.  -> the previous cap"""
    return int(lib().bg_two_ply_reply_sampling(int(cap), int(seed) & ((1 << 64) - 1)))


class HostPipeline:
    """The per-decision hot path for HOST-resident batches (bg_hostpipe_*): pinned (boards, players[, rolls]) in, (action, count) per
    item out.  Chunking, the copy / kernel overlap over n_streams streams and every device buffer live inside libbgarena.so; this class
    only holds the handle.  all_rolls=True: the inputs are POSITIONS, each expanded to the 21 rolls of DICE_ROLLS (items_per_chunk is
    then positions per chunk and the outputs have 21 entries per position).
    This is the call a CPU-side caller of get_all_possible_moves + generate_all_board_features + policy_network.forward +
    argmax/sample (reference worker.py:101-143) makes when its positions live in host memory."""

    def __init__(self, weights: PreparedWeights, items_per_chunk: int = 1 << 21, device=None, item_cap: int = 500, rows_per_item: int = 26,
                 n_streams: int = 3, all_rolls: bool = False):
        import ctypes as C

        self.dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.weights, self.all_rolls = weights, bool(all_rolls)
        self._h = C.c_void_p()
        check(lib().bg_hostpipe_create(C.byref(self._h), self.dev.index or 0, weights.H, int(items_per_chunk), int(self.all_rolls), int(item_cap),
                                       int(rows_per_item), max(1, int(n_streams))))

    def run(self, h_boards: torch.Tensor, h_players: torch.Tensor, h_rolls: Optional[torch.Tensor], h_actions: torch.Tensor, h_counts: torch.Tensor,
            temperature: float = 0.0, seed: int = 0) -> None:
        """Host tensors (pinned for asynchronous copies): [n,52] int8, [n] uint8, [n,2] uint8 (None with all_rolls), int32 outputs with
        one entry per item (21 per position with all_rolls).  Returns after enqueueing; the current stream is ordered after every
        chunk (synchronise it to read the results)."""
        n = h_boards.shape[0]
        per = 21 if self.all_rolls else 1
        for t, name in ((h_boards, "h_boards"), (h_players, "h_players"), (h_actions, "h_actions"), (h_counts, "h_counts")):
            if t.is_cuda or not t.is_contiguous():
                raise ValueError(f"{name} must be a contiguous host tensor")
        if h_actions.numel() < n * per or h_counts.numel() < n * per or h_actions.dtype != torch.int32 or h_counts.dtype != torch.int32:
            raise ValueError("h_actions / h_counts must be int32 with one entry per item")
        if h_boards.dtype != torch.int8 or h_players.dtype != torch.uint8 or h_players.numel() != n:
            raise ValueError("h_boards must be int8 [n,52] and h_players uint8 [n]")
        if not self.all_rolls and (h_rolls is None or h_rolls.is_cuda or h_rolls.dtype != torch.uint8 or h_rolls.numel() != 2 * n):
            raise ValueError("h_rolls must be a host uint8 [n,2] tensor")
        with torch.cuda.device(self.dev):
            check(lib().bg_hostpipe_run(self._h, h_boards.data_ptr(), h_players.data_ptr(), None if self.all_rolls else h_rolls.data_ptr(), n,
                                        self.weights.table.data_ptr(), float(temperature), seed & (2**64 - 1), h_actions.data_ptr(),
                                        h_counts.data_ptr(), torch.cuda.current_stream(self.dev).cuda_stream))

    def raise_for_status(self):
        """synchronises the pipe; raises if ANY chunk since the last call reported a capacity / invariant problem"""
        import ctypes as C

        st = C.c_int32(0)
        check(lib().bg_hostpipe_status(self._h, C.byref(st)))
        if st.value != 0:
            raise _lib.BgError(st.value, "HostPipeline: bg_movegen reported a capacity/invariant problem")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().bg_hostpipe_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
