"""ctypes binding for the CPU ORACLE (oracle/bg_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product (mlp-ppo-2ply-multi_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libbgoracle.so")

NFEAT = 198
BOARD_BYTES = 52
MAX_MOVES = 4096

DICE_ROLLS = [(a, b) for a in range(1, 7) for b in range(a, 7)]  # src/multi/two_ply.py:10-32 order


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (no GPU needed)."""
    src = os.path.join(_HERE, "bg_oracle.c")
    hdr = os.path.join(_HERE, "bg_oracle.h")
    if (
        not force
        and os.path.exists(_SO)
        and os.path.getmtime(_SO) >= max(os.path.getmtime(src), os.path.getmtime(hdr))
    ):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    base = ["gcc", "-O2", "-fPIC", "-std=c11", "-shared", "-o", _SO, src, "-lm"]
    try:
        subprocess.run(base[:1] + ["-fopenmp"] + base[1:], check=True, capture_output=True)
    except (subprocess.CalledProcessError, FileNotFoundError):
        subprocess.run(base, check=True)  # single-threaded fallback
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    L = C.CDLL(_SO)
    i8p = C.POINTER(C.c_int8)
    u8p = C.POINTER(C.c_uint8)
    f32p = C.POINTER(C.c_float)
    i64p = C.POINTER(C.c_int64)
    i32p = C.POINTER(C.c_int32)
    L.bgo_get_all_possible_moves.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    L.bgo_get_all_possible_moves.restype = C.c_int
    L.bgo_movegen_batch.argtypes = [i8p, u8p, u8p, C.c_int64, C.c_int64, i8p, u8p, i64p, C.c_int]
    L.bgo_movegen_batch.restype = C.c_int64
    L.bgo_encode_batch.argtypes = [i8p, u8p, C.c_int64, f32p, C.c_int]
    L.bgo_encode_batch.restype = None
    L.bgo_forward.argtypes = [f32p, C.c_int, f32p, C.c_int64, f32p]
    L.bgo_forward.restype = None
    L.bgo_eval_boards.argtypes = [f32p, C.c_int, i8p, u8p, C.c_int64, f32p]
    L.bgo_eval_boards.restype = None
    L.bgo_two_ply_batch.argtypes = [i8p, u8p, f32p, C.c_int64, f32p, C.c_int, C.c_int, C.c_float, C.c_float, f32p, i64p, C.c_int]
    L.bgo_two_ply_batch.restype = None
    L.bgo_env_create.argtypes = [C.c_int, u8p, C.c_int64, C.c_uint64]
    L.bgo_env_create.restype = C.c_void_p
    L.bgo_env_destroy.argtypes = [C.c_void_p]
    L.bgo_env_reset.argtypes = [C.c_void_p]
    L.bgo_env_step.argtypes = [C.c_void_p, C.c_int, f32p, C.POINTER(C.c_int)]
    L.bgo_env_step.restype = C.c_int
    for name in ("bgo_env_num_moves", "bgo_env_player", "bgo_env_win_type"):
        getattr(L, name).argtypes = [C.c_void_p]
        getattr(L, name).restype = C.c_int
    L.bgo_env_roll.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.bgo_env_board.argtypes = [C.c_void_p]
    L.bgo_env_board.restype = C.c_void_p
    L.bgo_env_afterstates.argtypes = [C.c_void_p]
    L.bgo_env_afterstates.restype = C.c_void_p
    L.bgo_env_moves.argtypes = [C.c_void_p]
    L.bgo_env_moves.restype = C.c_void_p
    L.bgo_env_tape_pos.argtypes = [C.c_void_p]
    L.bgo_env_tape_pos.restype = C.c_int64
    L.bgo_play_episode.argtypes = [C.c_void_p, f32p, C.c_int, C.c_float, C.POINTER(C.c_uint64), C.c_int, C.c_void_p,
                                   i32p, i32p, u8p, u8p, f32p, f32p, f32p, i8p]
    L.bgo_play_episode.restype = None
    L.bgo_selfplay_bench.argtypes = [f32p, C.c_int, C.c_float, C.c_int64, C.c_uint64, C.c_int, i64p, i64p]
    L.bgo_selfplay_bench.restype = C.c_int64
    L.bgo_movegen_eval_bench.argtypes = [i8p, u8p, u8p, C.c_int64, f32p, C.c_int, C.c_int, C.POINTER(C.c_double)]
    L.bgo_movegen_eval_bench.restype = C.c_int64
    L.bgo_random_positions.argtypes = [C.c_int64, C.c_uint64, i8p, u8p, C.c_int]
    L.bgo_random_positions.restype = None
    for name in ("bgo_check_game_over", "bgo_check_gammon", "bgo_check_backgammon", "bgo_made_five_prime", "bgo_is_closed_out"):
        getattr(L, name).argtypes = [C.c_void_p, C.c_int]
        getattr(L, name).restype = C.c_int
    _lib = L
    return L


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def initial_board() -> np.ndarray:
    b = np.zeros(52, np.int8)
    b[0], b[11], b[16], b[18] = 2, 5, 3, 5
    b[24 + 23], b[24 + 12], b[24 + 7], b[24 + 5] = 2, 5, 3, 5
    return b


def pack_weights(state_dict) -> np.ndarray:
    """state_dict (fc1.weight[H,198], fc1.bias[H], value_head.weight[1,H], value_head.bias[1]) ->
    packed fp32 [W1t(198,H) | b1 | w2 | b2]."""
    g = lambda k: np.asarray(state_dict[k].detach().cpu().numpy() if hasattr(state_dict[k], "detach") else state_dict[k], np.float32)
    W1 = g("fc1.weight")
    return np.concatenate([W1.T.reshape(-1), g("fc1.bias").reshape(-1), g("value_head.weight").reshape(-1), g("value_head.bias").reshape(-1)]).astype(np.float32)


def legal_moves(board: np.ndarray, player: int, roll) -> tuple[np.ndarray, np.ndarray]:
    """-> (boards int8[n,52], submoves uint8[n,4,3] (255 padded))"""
    boards = np.ascontiguousarray(board, np.int8).reshape(1, 52)
    off = np.zeros(2, np.int64)
    players = np.array([player], np.uint8)
    rolls = np.array([roll], np.uint8).reshape(1, 2)
    ob = np.zeros((MAX_MOVES, 52), np.int8)
    om = np.zeros((MAX_MOVES, 4, 3), np.uint8)
    n = lib().bgo_movegen_batch(_p(boards, C.c_int8), _p(players, C.c_uint8), _p(rolls, C.c_uint8), 1, MAX_MOVES,
                                _p(ob, C.c_int8), _p(om, C.c_uint8), _p(off, C.c_int64), 1)
    assert n >= 0
    return ob[:n].copy(), om[:n].copy()


def movegen_batch(boards, players, rolls, want_moves=True, nthreads=0):
    """CSR batch move generation. -> (offsets int64[B+1], out_boards int8[T,52], out_submoves uint8[T,4,3] | None)"""
    boards = np.ascontiguousarray(boards, np.int8).reshape(-1, 52)
    B = boards.shape[0]
    players = np.ascontiguousarray(players, np.uint8).reshape(B)
    rolls = np.ascontiguousarray(rolls, np.uint8).reshape(B, 2)
    off = np.zeros(B + 1, np.int64)
    total = lib().bgo_movegen_batch(_p(boards, C.c_int8), _p(players, C.c_uint8), _p(rolls, C.c_uint8), B, 0, None, None,
                                    _p(off, C.c_int64), nthreads)
    total = abs(total)
    ob = np.zeros((max(total, 1), 52), np.int8)
    om = np.zeros((max(total, 1), 4, 3), np.uint8) if want_moves else None
    t2 = lib().bgo_movegen_batch(_p(boards, C.c_int8), _p(players, C.c_uint8), _p(rolls, C.c_uint8), B, max(total, 1),
                                 _p(ob, C.c_int8), _p(om, C.c_uint8) if want_moves else None, _p(off, C.c_int64), nthreads)
    assert t2 == total
    return off, ob[:total], (om[:total] if want_moves else None)


def encode(boards, flags, nthreads=0) -> np.ndarray:
    boards = np.ascontiguousarray(boards, np.int8).reshape(-1, 52)
    n = boards.shape[0]
    flags = np.ascontiguousarray(flags, np.uint8).reshape(n)
    out = np.zeros((n, NFEAT), np.float32)
    lib().bgo_encode_batch(_p(boards, C.c_int8), _p(flags, C.c_uint8), n, _p(out, C.c_float), nthreads)
    return out


def forward(packed: np.ndarray, H: int, feats: np.ndarray) -> np.ndarray:
    feats = np.ascontiguousarray(feats, np.float32).reshape(-1, NFEAT)
    out = np.zeros(feats.shape[0], np.float32)
    packed = np.ascontiguousarray(packed, np.float32)
    lib().bgo_forward(_p(packed, C.c_float), H, _p(feats, C.c_float), feats.shape[0], _p(out, C.c_float))
    return out


def value(packed, H, boards, flags) -> np.ndarray:
    """double-accumulated oracle value of boards (features -> forward)."""
    return forward(packed, H, encode(boards, flags))


def two_ply(cand_boards, mover, S, packed, H, top_k=5, alpha=1.0, beta=0.9, nthreads=0):
    cand_boards = np.ascontiguousarray(cand_boards, np.int8).reshape(-1, 52)
    n = cand_boards.shape[0]
    mover = np.ascontiguousarray(mover, np.uint8).reshape(n)
    S = np.ascontiguousarray(S, np.float32).reshape(n)
    packed = np.ascontiguousarray(packed, np.float32)
    out = np.zeros(n, np.float32)
    nrep = np.zeros(n, np.int64)
    lib().bgo_two_ply_batch(_p(cand_boards, C.c_int8), _p(mover, C.c_uint8), _p(S, C.c_float), n, _p(packed, C.c_float), H,
                            top_k, alpha, beta, _p(out, C.c_float), _p(nrep, C.c_int64), nthreads)
    return out, nrep


class EpisodeStats(C.Structure):
    _fields_ = [("n_steps", C.c_int32), ("n_decisions", C.c_int32), ("n_passes", C.c_int32), ("win_type", C.c_int32),
                ("winner", C.c_int32), ("n_afterstates", C.c_int64), ("total_reward", C.c_float)]


class Env:
    """The reference's BackgammonEnv state machine (src/environments/backgammon_env.py) on a dice tape."""

    def __init__(self, tape=None, seed=0, max_legal_moves=500):
        self._tape = None if tape is None else np.ascontiguousarray(tape, np.uint8).reshape(-1, 2)
        self._e = lib().bgo_env_create(max_legal_moves, None if self._tape is None else _p(self._tape, C.c_uint8),
                                       0 if self._tape is None else self._tape.shape[0], seed)

    def __del__(self):
        if getattr(self, "_e", None):
            lib().bgo_env_destroy(self._e)
            self._e = None

    def reset(self):
        lib().bgo_env_reset(self._e)

    def step(self, action):
        r = C.c_float()
        info = C.c_int()
        done = lib().bgo_env_step(self._e, -1 if action is None else int(action), C.byref(r), C.byref(info))
        return r.value, bool(done), info.value

    @property
    def num_moves(self):
        return lib().bgo_env_num_moves(self._e)

    @property
    def player(self):
        return lib().bgo_env_player(self._e)

    @property
    def win_type(self):
        return lib().bgo_env_win_type(self._e)

    @property
    def roll(self):
        a, b = C.c_int(), C.c_int()
        lib().bgo_env_roll(self._e, C.byref(a), C.byref(b))
        return a.value, b.value

    @property
    def tape_pos(self):
        return lib().bgo_env_tape_pos(self._e)

    @property
    def board(self) -> np.ndarray:
        ptr = lib().bgo_env_board(self._e)
        return np.frombuffer((C.c_int8 * 52).from_address(ptr), np.int8).copy()

    @property
    def afterstates(self) -> np.ndarray:
        n = self.num_moves
        ptr = lib().bgo_env_afterstates(self._e)
        return np.frombuffer((C.c_int8 * (52 * max(n, 1))).from_address(ptr), np.int8).reshape(-1, 52)[:n].copy()

    @property
    def moves(self) -> np.ndarray:
        """uint8[n,13]: 4x(start,end,hit) + n_submoves"""
        n = self.num_moves
        ptr = lib().bgo_env_moves(self._e)
        return np.frombuffer((C.c_uint8 * (13 * max(n, 1))).from_address(ptr), np.uint8).reshape(-1, 13)[:n].copy()

    def play_episode(self, packed, H, temperature=0.0, rng_seed=1, max_steps=300):
        packed = np.ascontiguousarray(packed, np.float32)
        st = EpisodeStats()
        rng = C.c_uint64(rng_seed)
        tr = dict(
            nmoves=np.zeros(max_steps, np.int32), action=np.zeros(max_steps, np.int32), roll=np.zeros((max_steps, 2), np.uint8),
            player=np.zeros(max_steps, np.uint8), v=np.zeros(max_steps, np.float32), vnext=np.zeros(max_steps, np.float32),
            reward=np.zeros(max_steps, np.float32), after=np.zeros((max_steps, 52), np.int8),
        )
        lib().bgo_play_episode(self._e, _p(packed, C.c_float), H, temperature, C.byref(rng), max_steps, C.byref(st),
                               _p(tr["nmoves"], C.c_int32), _p(tr["action"], C.c_int32), _p(tr["roll"], C.c_uint8),
                               _p(tr["player"], C.c_uint8), _p(tr["v"], C.c_float), _p(tr["vnext"], C.c_float),
                               _p(tr["reward"], C.c_float), _p(tr["after"], C.c_int8))
        nd = st.n_decisions
        tr = {k: v[:nd] for k, v in tr.items()}
        stats = {f[0]: getattr(st, f[0]) for f in EpisodeStats._fields_}
        return stats, tr


def selfplay_bench(packed, H, temperature, n_games, seed=0, nthreads=0):
    packed = np.ascontiguousarray(packed, np.float32)
    steps, decs = C.c_int64(), C.c_int64()
    after = lib().bgo_selfplay_bench(_p(packed, C.c_float), H, temperature, n_games, seed, nthreads, C.byref(steps), C.byref(decs))
    return after, steps.value, decs.value


def movegen_eval_bench(boards, players, rolls, packed, H, nthreads=0):
    boards = np.ascontiguousarray(boards, np.int8).reshape(-1, 52)
    B = boards.shape[0]
    players = np.ascontiguousarray(players, np.uint8).reshape(B)
    rolls = np.ascontiguousarray(rolls, np.uint8).reshape(B, 2)
    packed = np.ascontiguousarray(packed, np.float32)
    vs = C.c_double()
    n = lib().bgo_movegen_eval_bench(_p(boards, C.c_int8), _p(players, C.c_uint8), _p(rolls, C.c_uint8), B, _p(packed, C.c_float), H,
                                     nthreads, C.byref(vs))
    return n, vs.value


def random_positions(n: int, seed: int = 2026, nthreads: int = 0):
    """SURVEY.md section 8(d) config-2 position set: one position sampled uniformly from the plies of each
    uniform-random-legal-move playout from the initial board.  -> (boards int8[n,52], players uint8[n])"""
    boards = np.zeros((n, 52), np.int8)
    players = np.zeros(n, np.uint8)
    lib().bgo_random_positions(n, seed, _p(boards, C.c_int8), _p(players, C.c_uint8), nthreads)
    return boards, players


def all_rolls_items(boards, players):
    """positions x the 21 unordered rolls (src/multi/two_ply.py:10-32 order) -> item arrays"""
    boards = np.ascontiguousarray(boards, np.int8).reshape(-1, 52)
    n = boards.shape[0]
    rolls = np.array(DICE_ROLLS, np.uint8)
    ib = np.repeat(boards, 21, axis=0)
    ip = np.repeat(np.asarray(players, np.uint8), 21)
    ir = np.tile(rolls, (n, 1))
    return ib, ip, ir


LEARNER_NMETRICS = 6


class Learner:
    """oracle restatement of Trainer.update (reference src/agents/trainer.py:48-166)"""

    def __init__(self, packed, H, lr=1e-3, gamma=0.99, grad_clip=1.0):
        L = lib()
        L.bgo_learner_create.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_float, C.c_float, C.c_float]
        L.bgo_learner_create.restype = C.c_void_p
        L.bgo_learner_destroy.argtypes = [C.c_void_p]
        L.bgo_learner_update.argtypes = [C.c_void_p, C.POINTER(C.c_int8), C.POINTER(C.c_uint8), C.POINTER(C.c_float),
                                         C.POINTER(C.c_int64), C.c_int64, C.POINTER(C.c_float)]
        L.bgo_learner_get.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int64)]
        packed = np.ascontiguousarray(packed, np.float32)
        self.H = H
        self.n_params = 200 * H + 1
        assert packed.size == self.n_params
        self._l = L.bgo_learner_create(_p(packed, C.c_float), H, lr, gamma, grad_clip)

    def __del__(self):
        try:
            if getattr(self, "_l", None):
                lib().bgo_learner_destroy(self._l)
                self._l = None
        except Exception:  # interpreter shutdown
            pass

    def update(self, obs_boards, obs_flags, reward, ep_offsets) -> np.ndarray:
        """-> per-episode metrics float32 [E, 6] (loss, mean |td|, clipped grad norm, mean V, reward sum, length)"""
        obs_boards = np.ascontiguousarray(obs_boards, np.int8).reshape(-1, 52)
        obs_flags = np.ascontiguousarray(obs_flags, np.uint8)
        reward = np.ascontiguousarray(reward, np.float32)
        ep_offsets = np.ascontiguousarray(ep_offsets, np.int64)
        E = ep_offsets.size - 1
        met = np.zeros((E, LEARNER_NMETRICS), np.float32)
        lib().bgo_learner_update(self._l, _p(obs_boards, C.c_int8), _p(obs_flags, C.c_uint8), _p(reward, C.c_float),
                                 _p(ep_offsets, C.c_int64), E, _p(met, C.c_float))
        return met

    def state(self):
        """-> (packed, exp_avg, exp_avg_sq, step)"""
        p, m, v = (np.zeros(self.n_params, np.float32) for _ in range(3))
        st = C.c_int64()
        lib().bgo_learner_get(self._l, _p(p, C.c_float), _p(m, C.c_float), _p(v, C.c_float), C.byref(st))
        return p, m, v, st.value


def selfplay_episodes(packed, H, n_games, temperature=1.5, seed=0, max_steps=300):
    """n_games oracle self-play episodes in the learner's CSR form:
    (obs_boards int8[N,52], obs_flags uint8[N], reward fp32[N], ep_offsets int64[E+1], win_types int32[E])."""
    obs, flg, rew, off, wins = [], [], [], [0], []
    for g in range(n_games):
        env = Env(seed=seed * 1000003 + g)
        st, tr = env.play_episode(packed, H, temperature=temperature, rng_seed=seed * 7919 + g + 1, max_steps=max_steps)
        n = len(tr["reward"])
        if n == 0:
            continue
        ob = np.empty((n, 52), np.int8)
        ob[0] = initial_board()
        ob[1:] = tr["after"][:-1]
        obs.append(ob)
        flg.append(tr["player"].astype(np.uint8))
        rew.append(tr["reward"].astype(np.float32))
        off.append(off[-1] + n)
        wins.append(st["win_type"])
    return np.concatenate(obs), np.concatenate(flg), np.concatenate(rew), np.array(off, np.int64), np.array(wins, np.int32)
