#!/usr/bin/env python
"""Golden vectors for the env paths that the random / greedy games of make_golden.py never reach, produced by RUNNING THE
UNMODIFIED PYTHON REFERENCE (/root/reference) in the build container:

  env_shaping.npz   40 games of the reference BackgammonEnv under a point-making / hitting policy: steps that pay the once-per-player
                    shaping rewards +0.30 (close-out) and +0.20 (five-prime), src/environments/backgammon_env.py:195-218,
                    env_helper.py:167-242.  Same record layout as env_random.npz.
  env_truncate.npz  positions with more than 500 legal moves through BackgammonEnv.update_legal_moves: the env keeps the FIRST
                    max_legal_moves = 500 (backgammon_env.py:35,262-272); the true count comes from get_all_possible_moves.
  worker_cap.npz    one game that does not finish within MAX_TIMESTEPS = 300 env steps through the unmodified
                    Worker.play_episode (src/multi/worker.py:78-174, cap at :101), under a stalling policy network: dice tape,
                    the actions the reference sampled, per-experience values / rewards, the 300-step cut.

    python tests/golden/make_golden_env_paths.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import reference_shim as shim  # noqa: E402

shim.install()
torch.set_num_threads(1)

from src.backgammon.moves.generate_all_moves import get_all_possible_moves  # noqa: E402
from src.backgammon.types import Player  # noqa: E402
from environments import BackgammonEnv, execute_full_move_on_board_copy  # noqa: E402
from agents import BackgammonPolicyNetwork  # noqa: E402
from multi.worker import Worker  # noqa: E402
from config import MAX_TIMESTEPS  # noqa: E402

H = 128


class TapeEnv(BackgammonEnv):
    """reference env whose dice (np.random) and step() arguments are recorded; nothing else is touched"""

    def __init__(self, *a, **k):
        self.tape, self.actions = [], []
        super().__init__(*a, **k)

    def roll_dice(self):
        super().roll_dice()
        self.tape.append(tuple(self.roll_result))

    def step(self, action):
        self.actions.append(-1 if action is None else int(action))
        return super().step(action)


def info_bits(info, done):
    return ((1 if "No legal" in str(info.get("info", "")) else 0) | (4 if info.get("close_out_reward") else 0) | (8 if info.get("prime_reward") else 0)
            | (16 if done else 0) | ({"regular": 1, "gammon": 2, "backgammon": 3}.get(info.get("win_type"), 0) << 8))


def point_making_action(env, rng):
    """maximise the mover's made points (feature 'count >= 2') + 3 x opponent checkers on the bar over the legal afterstates"""
    n = env.num_moves
    if rng.random() < 0.12:
        return int(rng.integers(0, n))
    f = env.legal_board_features[:n].numpy()
    me = int(env.current_player)
    made = f[:, me * 96 + 1:me * 96 + 96:4].sum(1)
    score = made + 3.0 * f[:, 192 + (1 - me)] * 2.0 + 1e-3 * rng.random(n)
    return int(np.argmax(score))


def gen_shaping(n_games=40):
    recs = dict(tape=[], tape_off=[0], action=[], reward=[], done=[], info=[], board=[], player=[], nmoves=[], roll=[], step_off=[0],
                start_player=[], start_nmoves=[], start_roll=[])
    rng = np.random.default_rng(4041)
    for g in range(n_games):
        np.random.seed(5000 + g)
        env = TapeEnv()
        env.tape = []
        env.reset()
        recs["start_player"].append(int(env.current_player))
        recs["start_nmoves"].append(env.num_moves)
        recs["start_roll"].append(tuple(env.roll_result))
        for _ in range(1200):
            a = -1 if env.num_moves == 0 else point_making_action(env, rng)
            _, r, done, info = env.step(None if a < 0 else a)
            recs["action"].append(a)
            recs["reward"].append(float(r))
            recs["done"].append(int(done))
            recs["info"].append(info_bits(info, done))
            recs["board"].append(shim.board_to_array(env.board))
            recs["player"].append(int(env.current_player))
            recs["nmoves"].append(env.num_moves)
            recs["roll"].append(tuple(env.roll_result))
            if done:
                break
        recs["tape"] += env.tape
        recs["tape_off"].append(len(recs["tape"]))
        recs["step_off"].append(len(recs["action"]))
    info = np.array(recs["info"], np.int32)
    np.savez_compressed(os.path.join(HERE, "env_shaping.npz"), tape=np.array(recs["tape"], np.uint8), tape_off=np.array(recs["tape_off"], np.int64),
                        action=np.array(recs["action"], np.int32), reward=np.array(recs["reward"], np.float32),
                        done=np.array(recs["done"], np.uint8), info=info, board=np.array(recs["board"], np.int8),
                        player=np.array(recs["player"], np.uint8), nmoves=np.array(recs["nmoves"], np.int32),
                        roll=np.array(recs["roll"], np.uint8), step_off=np.array(recs["step_off"], np.int64),
                        start_player=np.array(recs["start_player"], np.uint8), start_nmoves=np.array(recs["start_nmoves"], np.int32),
                        start_roll=np.array(recs["start_roll"], np.uint8))
    print("env_shaping:", len(recs["action"]), "steps;", int(((info & 4) != 0).sum()), "close-out rewards,", int(((info & 8) != 0).sum()),
          "prime rewards; reward values", sorted(set(np.round(recs["reward"], 2))), "; max legal moves", max(recs["nmoves"]))


def gen_truncate():
    """many distinct points x small doubles, opponent far away (the construction of tests/test_gpu_parity.py::test_movegen_capacity_tiers)"""
    rng = np.random.default_rng(3)
    out = dict(boards=[], players=[], rolls=[], true_count=[], kept=[], kept_off=[0])
    made = 0
    for k in range(96):
        npts = [8, 10, 12, 13, 15, 15][k % 6]
        b = np.zeros(52, np.int8)
        pts = rng.choice(np.arange(0, 20), size=npts, replace=False)
        for p in pts:
            b[p] += 1
        b[int(pts[0])] += 15 - b[:24].sum()
        b[24 + 23] = 15
        if npts not in (10, 12) or made >= 3:
            continue
        for roll in ([1, 1], [2, 2]):
            moves = get_all_possible_moves(Player(0), shim.array_to_board(b), list(roll))
            if not (500 < len(moves) <= 1500):
                continue
            env = BackgammonEnv()
            env.reset()
            env.board = shim.array_to_board_env(b)
            env.current_player = type(env.current_player)(0)
            env.roll_result = list(roll)
            env.update_legal_moves()
            assert env.num_moves == env.max_legal_moves == 500 and len(env.legal_moves) == 500 and float(env.action_mask.sum()) == 500
            kept = [shim.board_to_array(execute_full_move_on_board_copy(env.board, m)) for m in env.legal_moves]
            # the kept list is the head of the full list
            full_head = [shim.board_to_array(execute_full_move_on_board_copy(env.board, m)) for m in moves[:500]]
            assert np.array_equal(np.array(kept), np.array(full_head))
            out["boards"].append(b)
            out["players"].append(0)
            out["rolls"].append(roll)
            out["true_count"].append(len(moves))
            out["kept"] += kept
            out["kept_off"].append(len(out["kept"]))
            made += 1
            print("env_truncate: position", k, "roll", roll, "true count", len(moves), "-> kept 500", flush=True)
            break
    np.savez_compressed(os.path.join(HERE, "env_truncate.npz"), boards=np.array(out["boards"], np.int8), players=np.array(out["players"], np.uint8),
                        rolls=np.array(out["rolls"], np.uint8), true_count=np.array(out["true_count"], np.int32),
                        kept=np.array(out["kept"], np.int8), kept_off=np.array(out["kept_off"], np.int64))


def stalling_state_dict(hitw=4.32, backw=0.10, blotw=1.79, homew=1.54):
    """a value net whose greedy play hits, keeps checkers back and leaves blots: both sides keep sending each other to the bar"""
    W1 = np.zeros((H, 198), np.float32)
    w2 = np.zeros((1, H), np.float32)
    for u, (me, flag_opp, bar_opp) in enumerate([(0, 197, 193), (96, 196, 192)]):
        W1[u, bar_opp] = hitw
        for p in range(24):
            back = (23 - p) / 23.0 if me == 0 else p / 23.0
            W1[u, me + 4 * p + 0] += backw * back + blotw
            W1[u, me + 4 * p + 1] += -blotw
            if (p >= 18) if me == 0 else (p <= 5):
                W1[u, me + 4 * p + 0] += -homew
        W1[u, flag_opp] = -30.0
        w2[0, u] = 1.0
    return {"fc1.weight": torch.from_numpy(W1), "fc1.bias": torch.zeros(H), "value_head.weight": torch.from_numpy(w2), "value_head.bias": torch.zeros(1)}


def gen_worker_cap():
    sd = stalling_state_dict()
    for seed in range(20):
        np.random.seed(7000 + seed)
        torch.manual_seed(7000 + seed)
        w = object.__new__(Worker)  # Worker.__init__ only wires the ParameterManager; play_episode itself is unmodified
        w.worker_id = 0
        w.device = torch.device("cpu")
        w.temperature = 0.02
        w.policy_network = BackgammonPolicyNetwork()
        w.policy_network.load_state_dict(sd)
        env = TapeEnv(worker_id=0, device=w.device)
        env.tape, env.actions = [], []
        ep = w.play_episode(env)
        n_steps = len(env.actions)
        if n_steps >= MAX_TIMESTEPS and not env.game_over:
            break
    else:
        raise SystemExit("no capped game found")
    assert n_steps == MAX_TIMESTEPS == 300
    ex = ep.experiences
    packed = np.concatenate([sd["fc1.weight"].numpy().T.reshape(-1), sd["fc1.bias"].numpy(), sd["value_head.weight"].numpy().reshape(-1),
                             sd["value_head.bias"].numpy()]).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "worker_cap.npz"), packed=packed, H=np.int32(H), temperature=np.float32(w.temperature),
                        tape=np.array(env.tape, np.uint8), action=np.array(env.actions, np.int32), n_steps=np.int32(n_steps),
                        n_experiences=np.int32(len(ex)), state_value=np.array([e.state_value for e in ex], np.float32),
                        next_state_value=np.array([e.next_state_value for e in ex], np.float32),
                        reward=np.array([float(e.reward) for e in ex], np.float32), done=np.array([int(bool(e.done)) for e in ex], np.uint8),
                        final_board=shim.board_to_array(env.board), final_player=np.uint8(int(env.current_player)),
                        win_type=np.int32({"regular": 1, "gammon": 2, "backgammon": 3}.get(ep.win_type, 0)))
    print("worker_cap: seed", 7000 + seed, "steps", n_steps, "experiences", len(ex), "passes", sum(a < 0 for a in env.actions), "game over", env.game_over,
          "dice used", len(env.tape))


if __name__ == "__main__":
    gen_shaping()
    gen_worker_cap()
    gen_truncate()
