"""BackgammonEnv with the reference's reset/step contract and attributes (reference src/environments/backgammon_env.py:29-329),
for callers that drive one game at a time (Worker.play_episode, play_versus_ai).  Legal moves and the afterstate feature
matrix come from the CUDA kernels (bg_movegen + bg_encode, one launch per ply); dice are np.random.randint(1, 7) draws as in
the reference, so np.random.seed reproduces the same games.  For throughput use Arena (tens of thousands of games per GPU)."""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .board import ImmutableBoard, _device
from .types import Player, submoves_to_fullmove

REWARD_PASS = 0.0
REWARD_INVALID_ACTION = -1.0
REWARD_WIN_BACKGAMMON = 2.5
REWARD_WIN_GAMMON = 2.0
REWARD_WIN_NORMAL = 1.0
REWARD_CLOSE_OUT = 0.30
REWARD_MAKE_PRIME = 0.20


def get_opponent(player: Player) -> Player:
    return Player.PLAYER2 if player == Player.PLAYER1 else Player.PLAYER1


# terminal / shaping predicates (reference src/environments/env_helper.py:113-242) on the host-side board value
def check_game_over(board: ImmutableBoard, player: Player) -> bool:
    return board.borne_off[int(player)] >= 15


def check_for_gammon(board: ImmutableBoard, player: Player) -> bool:
    return board.borne_off[1 - int(player)] == 0


def check_for_backgammon(board: ImmutableBoard, player: Player) -> bool:
    opp = 1 - int(player)
    if board.borne_off[opp] > 0:
        return False
    pos = board.positions_1 if opp == 1 else board.positions_0
    home = range(18, 24) if player == Player.PLAYER1 else range(0, 6)
    return any(pos[i] > 0 for i in home) or board.bar[opp] > 0


def made_at_least_five_prime(board: ImmutableBoard, player: Player) -> bool:
    mine = board.positions_0 if player == Player.PLAYER1 else board.positions_1
    theirs = board.positions_1 if player == Player.PLAYER1 else board.positions_0
    order = range(24) if player == Player.PLAYER1 else range(23, -1, -1)
    run = 0
    for idx in order:
        run = run + 1 if mine[idx] >= 2 else 0
        if run >= 5:
            behind = range(idx + 1, 24) if player == Player.PLAYER1 else range(0, idx)
            if any(theirs[i] > 0 for i in behind):
                return True
    return False


def is_closed_out(board: ImmutableBoard, player: Player) -> bool:
    if board.bar[1 - int(player)] == 0:
        return False
    mine = board.positions_0 if player == Player.PLAYER1 else board.positions_1
    home = range(18, 24) if player == Player.PLAYER1 else range(0, 6)
    return all(mine[i] >= 2 for i in home)


class BackgammonEnv:
    metadata = {"render.modes": ["human"]}

    def __init__(self, worker_id=None, max_legal_moves=500, device=torch.device("cpu")):
        self.device = device
        self._cuda = _device()
        self.worker_id = worker_id
        self.max_legal_moves = max_legal_moves
        self.board_feature_length = 198
        self.board = ImmutableBoard.initial_board()
        self.current_player = Player.PLAYER1
        self.game_over = False
        self.win_type = None
        self.roll_result = None
        self.action_mask = torch.zeros(max_legal_moves, dtype=torch.float32, device=device)
        self.legal_board_features = torch.zeros((max_legal_moves, 198), dtype=torch.float32, device=device)
        self.legal_moves = []
        self._legal_boards = np.zeros((0, 52), np.int8)
        self.num_moves = 0
        self.previous_num_moves = 0
        self.close_out_reward_given = {Player.PLAYER1: False, Player.PLAYER2: False}
        self.prime_reward_given = {Player.PLAYER1: False, Player.PLAYER2: False}
        self.current_board_features = self._features(self.board, self.current_player)

    # ---- helpers --------------------------------------------------------------------------------------------------
    def _features(self, board, player):
        b = torch.from_numpy(board.to_array()).to(self._cuda).reshape(1, 52)
        f = torch.tensor([int(player)], dtype=torch.uint8, device=self._cuda)
        return ops.encode(b, f)[0].to(self.device)

    def roll_dice(self):
        self.roll_result = [np.random.randint(1, 7), np.random.randint(1, 7)]

    def get_observation(self):
        return self.current_board_features

    def set_board(self, new_board):
        self.board = new_board
        self.current_board_features = self._features(self.board, self.current_player)

    def set_current_player(self, new_player):
        self.current_player = new_player
        self.current_board_features = self._features(self.board, self.current_player)

    def pass_turn(self):
        self.set_current_player(get_opponent(self.current_player))

    def update_legal_moves(self):
        """movegen + afterstate encoding on the GPU (backgammon_env.py:223-305): first max_legal_moves entries are kept, the
        action mask is updated incrementally and rows of legal_board_features beyond num_moves stay stale, as in the reference."""
        dev = self._cuda
        b = torch.from_numpy(self.board.to_array()).to(dev).reshape(1, 52)
        p = torch.tensor([int(self.current_player)], dtype=torch.uint8, device=dev)
        r = torch.tensor([[int(self.roll_result[0]), int(self.roll_result[1])]], dtype=torch.uint8, device=dev)
        cap = self.max_legal_moves
        res = ops.movegen(b, p, r, item_cap=cap, pool_cap=max(cap, 1), want_submoves=True, want_owner=False)
        n = min(int(res.counts[0].item()), cap)
        off = int(res.offsets[0].item()) if n else 0
        boards = res.boards[off:off + n]
        self.legal_moves = [submoves_to_fullmove(m, self.current_player) for m in res.submoves[off:off + n].cpu().numpy()]
        self._legal_boards = boards.cpu().numpy()
        self.num_moves = n
        if n > self.previous_num_moves:
            self.action_mask[self.previous_num_moves:n].fill_(1.0)
        elif n < self.previous_num_moves:
            self.action_mask[n:self.previous_num_moves].zero_()
        self.previous_num_moves = n
        if n > 0:
            self.legal_board_features[:n].copy_(ops.encode(boards.contiguous(), res.flags[off:off + n].contiguous()).to(self.device))

    # ---- gym-style API ---------------------------------------------------------------------------------------------------
    def reset(self):
        self.set_board(ImmutableBoard.initial_board())
        self.game_over = False
        self.win_type = None
        self.roll_dice()
        while self.roll_result[0] == self.roll_result[1]:
            self.roll_dice()
        self.set_current_player(Player.PLAYER2 if self.roll_result[0] < self.roll_result[1] else Player.PLAYER1)
        self.roll_dice()
        while self.roll_result[0] == self.roll_result[1]:
            self.roll_dice()
        self.update_legal_moves()
        self.close_out_reward_given = {Player.PLAYER1: False, Player.PLAYER2: False}
        self.prime_reward_given = {Player.PLAYER1: False, Player.PLAYER2: False}
        return self.get_observation()

    def step(self, action):
        info = {"current_player": self.current_player}
        if self.game_over:
            return self.get_observation(), torch.tensor(0.0, device=self.device), True, info
        if self.num_moves == 0:
            self.pass_turn()
            self.roll_dice()
            self.update_legal_moves()
            return self.get_observation(), torch.tensor(REWARD_PASS, device=self.device), False, {**info, "info": "No legal actions, turn passed"}
        if action is None or not (-self.max_legal_moves <= action < self.max_legal_moves) or not self.action_mask[action].item():
            print(f"Worker {self.worker_id}: Invalid action {action}")
            return self.get_observation(), torch.tensor(REWARD_INVALID_ACTION, device=self.device), False, {**info, "info": "Invalid action"}
        self.set_board(ImmutableBoard.from_array(self._legal_boards[action], self.board.device))
        reward = torch.tensor(0.0, device=self.device)
        if check_game_over(self.board, self.current_player):
            if check_for_backgammon(self.board, self.current_player):
                reward, win_type = torch.tensor(REWARD_WIN_BACKGAMMON, device=self.device), "backgammon"
            elif check_for_gammon(self.board, self.current_player):
                reward, win_type = torch.tensor(REWARD_WIN_GAMMON, device=self.device), "gammon"
            else:
                reward, win_type = torch.tensor(REWARD_WIN_NORMAL, device=self.device), "regular"
            info.update({"winner": self.current_player, "win_type": win_type})
            self.win_type = win_type
            self.game_over = True
            done = True
        else:
            if is_closed_out(self.board, self.current_player) and not self.close_out_reward_given[self.current_player]:
                reward = reward + torch.tensor(REWARD_CLOSE_OUT, device=self.device)
                self.close_out_reward_given[self.current_player] = True
                info["close_out_reward"] = True
            if made_at_least_five_prime(self.board, self.current_player) and not self.prime_reward_given[self.current_player]:
                reward = reward + torch.tensor(REWARD_MAKE_PRIME, device=self.device)
                self.prime_reward_given[self.current_player] = True
                info["prime_reward"] = True
            done = False
            self.pass_turn()
            self.roll_dice()
            self.update_legal_moves()
        return self.get_observation(), reward, done, info
