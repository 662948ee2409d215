#!/usr/bin/env python
"""Generate golden vectors by RUNNING THE UNMODIFIED PYTHON REFERENCE (/root/reference).

Run in the build container only (the reference does not travel to the GPU box):
    python tests/golden/make_golden.py
Outputs (committed): tests/golden/{movegen,features,values,predicates,env_random,greedy_games,two_ply}.npz

Every array below is produced by reference code:
  get_all_possible_moves            src/backgammon/moves/generate_all_moves.py:7
  execute_full_move_on_board_copy   src/environments/env_helper.py:27
  ImmutableBoard.get_board_features src/backgammon/board/immutable_board.py:86
  BackgammonPolicyNetwork.forward   src/agents/policy_network.py:53
  BackgammonEnv.reset/step          src/environments/backgammon_env.py:92,130
  compute_weighted_opponent_response src/multi/two_ply.py:93 (random.sample disabled)
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
from environments import BackgammonEnv, execute_full_move_on_board_copy, generate_all_board_features  # noqa: E402
from agents import BackgammonPolicyNetwork  # noqa: E402
import multi.two_ply as two_ply  # noqa: E402

DICE_ROLLS = [tuple(r) for r in two_ply.DICE_ROLLS]
CKPT = os.path.join(shim.REF, "src", "play", "backgammon_256_standard_episode_2100000.pth")


class _NoSample:
    @staticmethod
    def sample(seq, k):
        return seq


def pack_weights(sd):
    W1 = sd["fc1.weight"].numpy().astype(np.float32)
    return np.concatenate([W1.T.reshape(-1), sd["fc1.bias"].numpy().reshape(-1), sd["value_head.weight"].numpy().reshape(-1),
                           sd["value_head.bias"].numpy().reshape(-1)]).astype(np.float32)


class TapeEnv(BackgammonEnv):
    """reference env whose dice are recorded (np.random) or replayed"""

    def __init__(self, *a, **k):
        self.tape = []
        super().__init__(*a, **k)

    def roll_dice(self):
        super().roll_dice()
        self.tape.append(tuple(self.roll_result))


def special_positions():
    z = lambda: np.zeros(52, np.int8)
    out = []
    b = shim.board_to_array(shim.array_to_board(z()).initial_board())
    out += [(b, 0), (b, 1)]
    # quirk Q1: P1 bar 1, point 10 x1, off 13; P2 6x2, 11x2, 23x11
    q = z()
    q[10] = 1
    q[48] = 1
    q[50] = 13
    q[24 + 6] = 2
    q[24 + 11] = 2
    q[24 + 23] = 11
    out += [(q, 0)]
    # bear-off races, with and without opposing anchors / blots in the home board
    r = z()
    r[18:24] = [3, 2, 2, 3, 3, 2]
    r[24:30] = [2, 3, 3, 2, 2, 3]
    out += [(r, 0), (r, 1)]
    r2 = z()
    r2[19] = 1
    r2[21] = 1
    r2[50] = 13
    r2[24 + 23] = 2
    r2[24 + 22] = 1
    r2[24 + 3] = 12
    out += [(r2, 0)]
    r3 = z()
    r3[24 + 4] = 1
    r3[24 + 2] = 2
    r3[24 + 0] = 1
    r3[51] = 11
    r3[0 + 1] = 1
    r3[0 + 3] = 2
    r3[22] = 12
    out += [(r3, 1)]
    # both on the bar, crowded entry
    c = z()
    c[48] = 2
    c[49] = 3
    c[0:6] = [2, 0, 2, 1, 0, 2]
    c[12] = 6
    c[24 + 18:24 + 24] = [2, 2, 0, 1, 2, 0]
    c[24 + 10] = 5
    out += [(c, 0), (c, 1)]
    # one checker left / game nearly over
    o = z()
    o[23] = 1
    o[50] = 14
    o[24 + 0] = 1
    o[51] = 14
    out += [(o, 0), (o, 1)]
    return out


def playout_positions(n_games, seed):
    """positions sampled from random-action games of the reference env"""
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    out = []
    env = TapeEnv()
    for _ in range(n_games):
        env.reset()
        traj = [(shim.board_to_array(env.board), int(env.current_player))]
        for _ in range(400):
            a = None if env.num_moves == 0 else int(rng.integers(0, env.num_moves))
            _, _, done, _ = env.step(a)
            if done:
                break
            traj.append((shim.board_to_array(env.board), int(env.current_player)))
        for k in rng.choice(len(traj), size=min(3, len(traj)), replace=False):
            out.append(traj[int(k)])
    return out


def gen_movegen(positions):
    boards, players, rolls, offsets, ob, om = [], [], [], [0], [], []
    for b, p in positions:
        ib = shim.array_to_board(b)
        for r in DICE_ROLLS:
            moves = get_all_possible_moves(Player(p), ib, list(r))
            boards.append(b)
            players.append(p)
            rolls.append(r)
            offsets.append(offsets[-1] + len(moves))
            om.append(shim.moves_to_array(moves))
            benv = shim.array_to_board_env(b)
            for m in moves:
                ob.append(shim.board_to_array(execute_full_move_on_board_copy(benv, m)))
    return dict(boards=np.array(boards, np.int8), players=np.array(players, np.uint8), rolls=np.array(rolls, np.uint8),
                offsets=np.array(offsets, np.int64), out_boards=np.array(ob, np.int8).reshape(-1, 52),
                out_submoves=np.concatenate(om, 0).astype(np.uint8))


def main():
    sd = torch.load(CKPT, map_location="cpu")
    net = BackgammonPolicyNetwork()
    net.load_state_dict(sd)
    packed = pack_weights(sd)

    # ---- movegen ----------------------------------------------------------------------------------
    positions = special_positions() + playout_positions(60, seed=11)
    mg = gen_movegen(positions)
    np.savez_compressed(os.path.join(HERE, "movegen.npz"), **mg)
    print("movegen:", len(mg["boards"]), "items", mg["offsets"][-1], "afterstates", "max", np.diff(mg["offsets"]).max())

    # ---- features + values -----------------------------------------------------------------------------
    sel = np.random.default_rng(5).choice(len(mg["out_boards"]), size=1500, replace=False)
    fb = mg["out_boards"][sel]
    fflag = np.random.default_rng(6).integers(0, 2, size=len(fb)).astype(np.uint8)
    feats = np.stack([shim.array_to_board(b).get_board_features(Player(int(f))).numpy() for b, f in zip(fb, fflag)])
    with torch.no_grad():
        vals = net.forward(torch.from_numpy(feats)).numpy()
    np.savez_compressed(os.path.join(HERE, "features.npz"), boards=fb, flags=fflag, features=feats.astype(np.float32))
    torch.manual_seed(0)
    net0 = BackgammonPolicyNetwork()  # Xavier init, seed 0 (config 1)
    with torch.no_grad():
        vals0 = net0.forward(torch.from_numpy(feats)).numpy()
    np.savez_compressed(os.path.join(HERE, "values.npz"), boards=fb, flags=fflag, packed=packed, values=vals.astype(np.float32),
                        packed_init0=pack_weights(net0.state_dict()), values_init0=vals0.astype(np.float32), H=np.int32(128))
    print("features/values:", feats.shape, float(vals.min()), float(vals.max()))

    # ---- terminal / shaping predicates (env_helper.py:113-242) ---------------------------------------------------
    from environments.env_helper import (check_for_backgammon, check_for_gammon, check_game_over, is_closed_out,
                                         made_at_least_five_prime)
    from backgammon.types.moves import Player as EPlayer

    prng = np.random.default_rng(21)
    pb = [mg["out_boards"][i] for i in prng.choice(len(mg["out_boards"]), size=1200, replace=False)]
    for _ in range(400):  # crafted: random runs of made points (primes / closed boards), bar + stragglers
        z = np.zeros(52, np.int8)
        pl = int(prng.integers(0, 2))
        start = int(prng.integers(0, 20))
        ln = int(prng.integers(3, 7))
        for i in range(start, min(24, start + ln)):
            z[pl * 24 + i] = int(prng.integers(1, 4)) if prng.random() < 0.15 else int(prng.integers(2, 4))
        if prng.random() < 0.5:
            lo = 18 if pl == 0 else 0
            z[pl * 24 + lo:pl * 24 + lo + 6] = prng.integers(1, 4, size=6) if prng.random() < 0.3 else prng.integers(2, 4, size=6)
        op = 1 - pl
        for _k in range(int(prng.integers(0, 4))):
            z[op * 24 + int(prng.integers(0, 24))] += 1
        z[48 + op] = int(prng.integers(0, 3))
        z[50 + op] = int(prng.integers(0, 2)) * int(prng.integers(0, 16))
        z[50 + pl] = int(prng.choice([0, 3, 15]))
        pb.append(z)
    pb = np.array(pb, np.int8)
    pred = np.zeros((len(pb), 2, 5), np.uint8)
    for i, b in enumerate(pb):
        ib = shim.array_to_board_env(b)
        for pl in (0, 1):
            P = EPlayer(pl)
            pred[i, pl] = [check_game_over(ib, P), check_for_gammon(ib, P), check_for_backgammon(ib, P),
                           made_at_least_five_prime(ib, P), is_closed_out(ib, P)]
    np.savez_compressed(os.path.join(HERE, "predicates.npz"), boards=pb, pred=pred)
    print("predicates:", pred.reshape(-1, 5).sum(0))

    # ---- env: random-action games with recorded dice ---------------------------------------------------------
    recs = dict(tape=[], tape_off=[0], action=[], reward=[], done=[], info=[], board=[], player=[], nmoves=[], roll=[], step_off=[0],
                start_player=[], start_nmoves=[], start_roll=[])
    rng = np.random.default_rng(77)
    for g in range(12):
        np.random.seed(1000 + g)
        env = TapeEnv()
        env.tape = []
        env.reset()
        recs["start_player"].append(int(env.current_player))
        recs["start_nmoves"].append(env.num_moves)
        recs["start_roll"].append(tuple(env.roll_result))
        for _ in range(600):
            a = -1 if env.num_moves == 0 else int(rng.integers(0, env.num_moves))
            _, r, done, info = env.step(None if a < 0 else a)
            bits = (1 if "No legal" in str(info.get("info", "")) else 0) | (4 if info.get("close_out_reward") else 0) | (
                8 if info.get("prime_reward") else 0) | (16 if done else 0)
            recs["action"].append(a)
            recs["reward"].append(float(r))
            recs["done"].append(int(done))
            recs["info"].append(bits | ({"regular": 1, "gammon": 2, "backgammon": 3}.get(info.get("win_type"), 0) << 8))
            recs["board"].append(shim.board_to_array(env.board))
            recs["player"].append(int(env.current_player))
            recs["nmoves"].append(env.num_moves)
            recs["roll"].append(tuple(env.roll_result))
            if done:
                break
        recs["tape"] += env.tape
        recs["tape_off"].append(len(recs["tape"]))
        recs["step_off"].append(len(recs["action"]))
    np.savez_compressed(os.path.join(HERE, "env_random.npz"), tape=np.array(recs["tape"], np.uint8), tape_off=np.array(recs["tape_off"], np.int64),
                        action=np.array(recs["action"], np.int32), reward=np.array(recs["reward"], np.float32),
                        done=np.array(recs["done"], np.uint8), info=np.array(recs["info"], np.int32), board=np.array(recs["board"], np.int8),
                        player=np.array(recs["player"], np.uint8), nmoves=np.array(recs["nmoves"], np.int32),
                        roll=np.array(recs["roll"], np.uint8), step_off=np.array(recs["step_off"], np.int64),
                        start_player=np.array(recs["start_player"], np.uint8), start_nmoves=np.array(recs["start_nmoves"], np.int32),
                        start_roll=np.array(recs["start_roll"], np.uint8))
    print("env_random:", len(recs["action"]), "steps", "rewards", sorted(set(np.round(recs["reward"], 2))))

    # ---- greedy games (worker.py:78-174 loop with argmax, appendix B) ----------------------------------------
    g = dict(tape=[], tape_off=[0], dec_off=[0], nmoves=[], action=[], roll=[], player=[], v=[], vnext=[], reward=[], gap=[], after=[],
             n_steps=[], n_passes=[], win_type=[], winner=[])
    for s in range(6):
        np.random.seed(s)
        env = TapeEnv()
        env.tape = []
        obs = env.reset()
        done, steps, passes = False, 0, 0
        while not done and steps < 300:
            n = env.num_moves
            if n == 0:
                obs, _, done, _ = env.step(None)
                steps += 1
                passes += 1
                continue
            x = torch.cat([obs.unsqueeze(0), env.legal_board_features[:n]], 0)
            with torch.no_grad():
                v = net.forward(x)
            a = int(torch.argmax(v[1:]).item())
            srt = torch.sort(v[1:], descending=True)[0]
            g["gap"].append(float(srt[0] - srt[1]) if n > 1 else 1e9)
            g["nmoves"].append(n)
            g["action"].append(a)
            g["roll"].append(tuple(env.roll_result))
            g["player"].append(int(env.current_player))
            g["v"].append(float(v[0]))
            g["vnext"].append(float(v[1 + a]))
            mover = env.current_player
            obs, r, done, info = env.step(a)
            g["reward"].append(float(r))
            g["after"].append(shim.board_to_array(env.board))
            steps += 1
        g["tape"] += env.tape
        g["tape_off"].append(len(g["tape"]))
        g["dec_off"].append(len(g["action"]))
        g["n_steps"].append(steps)
        g["n_passes"].append(passes)
        g["win_type"].append({"regular": 1, "gammon": 2, "backgammon": 3}.get(env.win_type, 0))
        g["winner"].append(int(mover))
        print("greedy seed", s, "steps", steps, "passes", passes, env.win_type, "winner P%d" % (int(mover) + 1))
    np.savez_compressed(os.path.join(HERE, "greedy_games.npz"), tape=np.array(g["tape"], np.uint8), tape_off=np.array(g["tape_off"], np.int64),
                        dec_off=np.array(g["dec_off"], np.int64), nmoves=np.array(g["nmoves"], np.int32), action=np.array(g["action"], np.int32),
                        roll=np.array(g["roll"], np.uint8), player=np.array(g["player"], np.uint8), v=np.array(g["v"], np.float32),
                        vnext=np.array(g["vnext"], np.float32), reward=np.array(g["reward"], np.float32), gap=np.array(g["gap"], np.float32),
                        after=np.array(g["after"], np.int8), n_steps=np.array(g["n_steps"], np.int32), n_passes=np.array(g["n_passes"], np.int32),
                        win_type=np.array(g["win_type"], np.int32), winner=np.array(g["winner"], np.int32))

    # ---- 2-ply ------------------------------------------------------------------------------------------------
    two_ply.random = _NoSample  # harness deviation: no 50-reply subsampling
    rng = np.random.default_rng(9)
    cand_idx = rng.choice(len(mg["out_boards"]), size=24, replace=False)
    # owner item of each afterstate -> mover
    owner = np.searchsorted(mg["offsets"], cand_idx, side="right") - 1
    cb = mg["out_boards"][cand_idx]
    mover = mg["players"][owner]
    W5, W1 = [], []
    for b, m in zip(cb, mover):
        ib = shim.array_to_board(b)
        W5.append(two_ply.compute_weighted_opponent_response(ib, Player(1 - int(m)), net))
    feats_c = np.stack([shim.array_to_board(b).get_board_features(Player(int(m))).numpy() for b, m in zip(cb, mover)])
    with torch.no_grad():
        S = net.forward(torch.from_numpy(feats_c)).numpy()
    np.savez_compressed(os.path.join(HERE, "two_ply.npz"), cand_boards=cb, mover=mover.astype(np.uint8), S=S.astype(np.float32),
                        W_top5=np.array(W5, np.float64), score=(1.0 * S.astype(np.float64) - 0.9 * np.array(W5)))
    print("two_ply:", len(cb), "W range", min(W5), max(W5))


if __name__ == "__main__":
    main()
