"""Build-container only: games/s of the UNMODIFIED Python reference (Worker.play_episode, one process) next to the C oracle port on the same core.
The GPU box has no /root/reference, so bench.py times the port there; this gives the factor between the two."""
import sys, time, os
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from oracle import reference_shim as shim, pyoracle as po
shim.install()
torch.set_num_threads(1)
from multi.worker import Worker
from environments import BackgammonEnv
from agents import BackgammonPolicyNetwork

class PM:
    def __init__(self, sd): self.sd = sd
    def get_parameters(self, device=None): return self.sd
    def get_version(self): return 1
    def get_temperature(self): return 1.5
torch.manual_seed(0)
net = BackgammonPolicyNetwork()
w = Worker.__new__(Worker)
w.worker_id = 0
w.policy_network = net
w.temperature = 1.5
np.random.seed(0)
env = BackgammonEnv()
n_games, t0, steps, after = 12, time.perf_counter(), 0, 0
import io, contextlib
for g in range(n_games):
    with contextlib.redirect_stdout(io.StringIO()):
        ep = w.play_episode(env)
    steps += len(ep.experiences)
dt = time.perf_counter() - t0
print(f"python reference (unmodified Worker.play_episode, 1 process, torch 1 thread): {n_games / dt:.3f} games/s, {steps / dt:.1f} decisions/s over {n_games} games ({dt:.1f} s)")
packed = po.pack_weights(net.state_dict())
t0 = time.perf_counter()
a, s, d = po.selfplay_bench(packed, 128, 1.5, 2000, seed=0, nthreads=1)
dt2 = time.perf_counter() - t0
print(f"C oracle port (bgo_selfplay_bench, 1 thread): {2000 / dt2:.1f} games/s, {d / dt2:.0f} decisions/s, {a / dt2:.0f} afterstates/s")
print(f"ratio port / reference per core: {(2000 / dt2) / (n_games / dt):.0f}x")
