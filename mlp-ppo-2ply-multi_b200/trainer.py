"""Trainer with the reference's surface (reference src/agents/trainer.py:11-228): Trainer(parameter_manager, device).update(episodes)
with len(episodes) == 200 else ValueError; per episode IN ORDER one forward pass, TD(0) targets, mse loss, backward,
clip_grad_norm_(1.0), one Adam step; then parameter_manager.set_parameters(...).

What changes: the 200 sequential forward/backward/Adam steps run inside ONE kernel launch (bg_learner_update, csrc/learner.cu: a
persistent thread-block cluster with the optimiser state in shared memory) instead of ~200 x 30 PyTorch launches, and the episodes
can be consumed straight from the arena's compact records (`update(EpisodeBatch)`) without ever building the [T,198] tensors.
There is no PyTorch/CPU fallback: without libbgarena.so and a CUDA device this raises."""
from __future__ import annotations

import ctypes as C
import time
from typing import Optional

import torch

from . import ops
from ._lib import check, lib
from .episode import EpisodeBatch

# reference src/config/configuration.py:7,17-22
MIN_EPISODES_TO_TRAIN = 200
GAMMA = 0.99
LEARNING_RATE = 1e-3
GRAD_CLIP_THRESHOLD = 1.0
LR_DECAY = 0.99
LR_DECAY_STEPS = 100_000

NMETRICS = 6
MAX_T = 320


def features_to_boards(obs: torch.Tensor):
    """Exact inverse of the 198-feature encoding (reference immutable_board.py:86-128): fp32 [N,198] -> (int8 [N,52] boards,
    uint8 [N] flag).  Lets reference-format Episodes (whose observations are feature tensors) feed the board-based learner."""
    obs = obs.to(torch.float32)
    pts = obs[:, :192].reshape(-1, 48, 4)
    cnt = pts[..., 0] + pts[..., 1] + pts[..., 2] + 2.0 * pts[..., 3]
    tail = torch.stack([2.0 * obs[:, 192], 2.0 * obs[:, 194], 15.0 * obs[:, 193], 15.0 * obs[:, 195]], dim=1)  # bar0, bar1, off0, off1
    boards = torch.cat([cnt, tail], dim=1).round().to(torch.int8)
    return boards.contiguous(), (obs[:, 197] > 0.5).to(torch.uint8).contiguous()


class TD0Learner:
    """Thin owner of a bg_learner handle (include/bgarena.h): packed weights + Adam moments resident on one GPU."""

    def __init__(self, hidden_size: int = 128, device=None, lr: float = LEARNING_RATE, gamma: float = GAMMA, grad_clip: Optional[float] = GRAD_CLIP_THRESHOLD):
        if not torch.cuda.is_available():
            raise RuntimeError("TD0Learner needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.device.type != "cuda":
            raise RuntimeError("TD0Learner needs a CUDA device (there is no CPU fallback)")
        self.H = int(hidden_size)
        self.n_params = 200 * self.H + 1
        self._h = C.c_void_p()
        check(lib().bg_learner_create(C.byref(self._h), self.device.index or 0, self.H, float(lr), float(gamma),
                                      float(grad_clip) if grad_clip is not None else 0.0))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().bg_learner_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def set_parameters(self, weights, reset_optimizer: bool = False):
        packed = weights if isinstance(weights, torch.Tensor) else ops.pack_weights(weights)
        packed = packed.to(self.device, torch.float32).contiguous()
        if packed.numel() != self.n_params:
            raise ValueError(f"weights do not match hidden_size={self.H}")
        check(lib().bg_learner_set_parameters(self._h, packed.data_ptr(), int(reset_optimizer), self._stream()))
        self._keep = packed  # alive until the copy has run

    def packed(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if out is None:
            out = torch.empty(self.n_params, dtype=torch.float32, device=self.device)
        check(lib().bg_learner_get_parameters(self._h, out.data_ptr(), self._stream()))
        return out

    def state_dict(self) -> dict:
        return ops.unpack_weights(self.packed(), self.H)

    def optimizer_state(self):
        """(exp_avg, exp_avg_sq) in packed order and the Adam step count."""
        m = torch.empty(self.n_params, dtype=torch.float32, device=self.device)
        v = torch.empty_like(m)
        step = torch.zeros(1, dtype=torch.int64, device=self.device)
        check(lib().bg_learner_get_optimizer(self._h, m.data_ptr(), v.data_ptr(), step.data_ptr(), self._stream()))
        return m, v, int(step.item())

    def update(self, boards: torch.Tensor, flags: torch.Tensor, reward: torch.Tensor, ep_offsets: torch.Tensor, n_episodes: Optional[int] = None,
               records: bool = False, check_status: bool = True, ep_len: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One sequential pass over the episodes (CSR).  records=False: boards/flags are the observations; records=True: they are the
        arena's after_boards/meta as drained.  -> per-episode metrics fp32 [E,6] (device): loss, mean |TD|, clipped grad norm, mean V,
        reward sum, length."""
        boards = ops._req(boards, torch.int8, "boards").reshape(-1, 52)
        flags = ops._req(flags, torch.uint8, "flags")
        reward = ops._req(reward, torch.float32, "reward")
        ep_offsets = ops._req(ep_offsets, torch.int64, "ep_offsets")
        E = ep_offsets.numel() - 1 if n_episodes is None else int(n_episodes)
        if ep_len is not None:
            ep_len = ops._req(ep_len, torch.int32, "ep_len")
        met = torch.zeros((max(E, 0), NMETRICS), dtype=torch.float32, device=self.device)
        status = torch.zeros(1, dtype=torch.int32, device=self.device)
        check(lib().bg_learner_update(self._h, boards.data_ptr(), flags.data_ptr(), reward.data_ptr(), ep_offsets.data_ptr(), ops._ptr(ep_len), E, int(records),
                                      met.data_ptr(), status.data_ptr(), self._stream()))
        self.last_status = status  # device int32[1]; read lazily by callers that must not synchronise here
        if check_status and int(status.item()) != 0:
            raise RuntimeError(f"bg_learner_update: an episode exceeded {MAX_T} experiences and was skipped")
        return met

    def update_batch(self, batch: EpisodeBatch, check_status: bool = True) -> torch.Tensor:
        """Consume drained arena episodes without materialising observations (zero-copy hand-off)."""
        return self.update(batch.after_boards, batch.meta, batch.reward, batch.ep_offsets, n_episodes=batch.n_episodes, records=True,
                           check_status=check_status, ep_len=batch.ep_len)


class _LearnerNetworkView:
    """trainer.policy_network: state_dict() / load_state_dict() of the weights held by the CUDA learner (reference trainer.py:21-24)."""

    def __init__(self, trainer):
        self._t = trainer

    def state_dict(self):
        return ops.unpack_weights(self._t.learner.packed(), self._t.learner.H)

    def load_state_dict(self, state_dict, strict: bool = True):
        self._t.learner.set_parameters(state_dict, reset_optimizer=False)

    def __call__(self, x):
        from .policy_network import BackgammonPolicyNetwork

        net = BackgammonPolicyNetwork(hidden_size=self._t.learner.H)
        net.load_state_dict({k: v.cpu() for k, v in self.state_dict().items()})
        return net(x.cpu())

    forward = __call__


class Trainer:
    """Drop-in for the reference Trainer (src/agents/trainer.py).  `episodes` may be a list of Episode objects in the reference's
    format (observations = [198] tensors) or an EpisodeBatch drained from the Arena."""

    def __init__(self, parameter_manager, device=None, s3_bucket_name=None, s3_log_prefix="logs/", logger=None):
        self.parameter_manager = parameter_manager
        self.device = torch.device(device) if device is not None else torch.device(f"cuda:{torch.cuda.current_device()}")
        state_dict = self.parameter_manager.get_parameters()
        H = state_dict["fc1.weight"].shape[0]
        self.learner = TD0Learner(H, self.device, lr=LEARNING_RATE, gamma=GAMMA, grad_clip=GRAD_CLIP_THRESHOLD)
        self.learner.set_parameters(state_dict, reset_optimizer=True)
        self.gamma = GAMMA
        self.total_episodes = 0
        self.grad_clip = GRAD_CLIP_THRESHOLD
        self.lr_decay = LR_DECAY  # carried but unused, as in the reference
        self.lr_decay_steps = LR_DECAY_STEPS
        self.batch_episode_size = MIN_EPISODES_TO_TRAIN
        self.logger = logger  # optional object with add_scalar/add_scalars (the reference's S3Logger surface); S3 is out of scope
        self.last_metrics: dict = {}
        self._pending = None

    @property
    def policy_network(self):
        """The reference exposes the network being trained (main.py:101 calls trainer.policy_network.load_state_dict(state_dict)); here the
        weights live in the CUDA learner, so this is a view with the same two methods."""
        return _LearnerNetworkView(self)

    def update(self, episodes):
        """Reference semantics (trainer.py:48-228): train on exactly 200 episodes, hand the new weights to the parameter manager,
        return (and log) the batch metrics.  Blocks until the update has run."""
        self.update_async(episodes)
        return self.finish()

    def update_async(self, episodes):
        """Enqueue the update on the current CUDA stream without any host synchronisation; finish() publishes and reports.  Lets
        the caller overlap the (8-SM) learner kernel with arena self-play on another stream."""
        n = episodes.n_episodes if isinstance(episodes, EpisodeBatch) else len(episodes)
        if n != self.batch_episode_size:
            raise ValueError(f"Expected {self.batch_episode_size} episodes, but got {n}.")
        if not isinstance(episodes, EpisodeBatch) and any(len(ep.experiences) > MAX_T for ep in episodes):
            raise ValueError(f"an episode has more than {MAX_T} experiences (the learner's limit; the reference caps episodes at 300 steps)")
        if self._pending is not None:
            self.finish()
        start = time.time()
        self.total_episodes += n
        dev = self.device
        if isinstance(episodes, EpisodeBatch):
            met = self.learner.update_batch(episodes, check_status=False)
            status = self.learner.last_status

            def summary():
                # built only when finish() is asked for the metrics (~25 small device ops: a training loop that publishes every update
                # but reads metrics now and then should not pay their launch cost per update); `episodes` is kept alive and must stay
                # unchanged until then
                info = episodes.ep_info[:n].to(torch.float32)
                lens = episodes.episode_lengths().to(torch.float32)
                # padded batches (all_gather_episodes(compact=False)) carry zero-length filler episodes: averages are over the real ones
                n_real = (lens > 0).sum().clamp(min=1).to(torch.float32)
                wins = torch.stack([(info[:, 0] == k).sum() for k in (1, 2, 3)]).to(torch.float32)
                seen = torch.stack([((episodes.ep_info[:n, 8] >> p) & 1).sum() for p in (0, 1)]).to(torch.float32)
                # the reference adds the episode's count dict once per EXPERIENCE (trainer.py:88-100)
                close = torch.stack([(info[:, 4 + p] * lens).sum() for p in (0, 1)])
                prime = torch.stack([(info[:, 6 + p] * lens).sum() for p in (0, 1)])
                return torch.cat([met.sum(dim=0) / n_real, wins, seen, close, prime, status.to(torch.float32)])  # read back once, in finish()

            host = None
        else:
            obs = torch.stack([x.observation for ep in episodes for x in ep.experiences]).to(dev)
            rew = torch.stack([torch.as_tensor(x.reward, dtype=torch.float32).reshape(()) for ep in episodes for x in ep.experiences]).to(dev)
            off = torch.tensor([0] + [len(ep.experiences) for ep in episodes], dtype=torch.int64).cumsum(0).to(dev)
            boards, flags = features_to_boards(obs)
            met = self.learner.update(boards, flags, rew.contiguous(), off, check_status=False)
            row = torch.cat([met.mean(dim=0), self.learner.last_status.to(torch.float32)])
            summary = lambda: row  # noqa: E731
            wins = {"regular": 0, "gammon": 0, "backgammon": 0}
            close, prime = {}, {}
            for ep in episodes:
                if ep.win_type in wins:
                    wins[ep.win_type] += 1
                for pid, c in ep.close_out_counts.items():
                    close[pid] = close.get(pid, 0) + c * len(ep.experiences)
                for pid, c in ep.prime_reward_counts.items():
                    prime[pid] = prime.get(pid, 0) + c * len(ep.experiences)
            host = (wins, close, prime)
        packed = self.learner.packed()
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(dev))
        self._pending = (summary, host, packed, done, start, episodes)  # `episodes` kept alive until the kernel has consumed it

    def finish(self, metrics: bool = True):
        """Wait (stream-ordered) for the pending update, hand its weights to the parameter manager (trainer.py:166) and return the
        metrics the reference logs (trainer.py:195-228).  metrics=False skips the read-back of the metrics (no host synchronisation:
        publication is stream-ordered); the learner's status is then checked by the next finish() that reads."""
        if self._pending is None:
            return self.last_metrics
        summary, host, packed, done, start, _ = self._pending
        self._pending = None
        torch.cuda.current_stream(self.device).wait_event(done)
        vals = None
        if metrics:
            vals = summary().tolist()
            if vals[-1] != 0:  # checked BEFORE the weights are published: a partial update never reaches the arenas
                raise RuntimeError(f"bg_learner_update: an episode exceeded {MAX_T} experiences and was skipped")
        if hasattr(self.parameter_manager, "set_packed"):
            self.parameter_manager.set_packed(packed, self.learner.H)  # device blob: no host round trip, one broadcast when distributed
        else:
            self.parameter_manager.set_parameters(ops.unpack_weights(packed, self.learner.H))
        if vals is None:
            return self.last_metrics
        avg = vals[:6]  # trainer.py:157-163: totals / batch_size
        if host is None:
            wins = dict(zip(("regular", "gammon", "backgammon"), (int(x) for x in vals[6:9])))
            close = {p: int(vals[11 + p]) for p in (0, 1) if vals[9 + p] > 0}
            prime = {p: int(vals[13 + p]) for p in (0, 1) if vals[9 + p] > 0}
        else:
            wins, close, prime = host
        self.last_metrics = {
            "Loss/Training Loss": avg[0], "TD Error/Mean TD Error": avg[1], "Gradients/Gradient Norm": avg[2],
            "Values/Average Predicted Value": avg[3], "Rewards/Average Reward per Episode": avg[4],
            "Episode/Average Episode Length": avg[5], "Wins": wins, "close_out_counts": close, "prime_reward_counts": prime,
            "update_seconds": time.time() - start,
        }
        if self.logger is not None:
            step = self.total_episodes
            for tag in list(self.last_metrics)[:6]:
                self.logger.add_scalar(tag, self.last_metrics[tag], step)
            for pid, c in close.items():
                self.logger.add_scalar(f"Rewards/CloseOutReward_Player{pid}", c, step)
            for pid, c in prime.items():
                self.logger.add_scalar(f"Rewards/PrimeReward_Player{pid}", c, step)
            self.logger.add_scalars("Wins", wins, step)
        return self.last_metrics
