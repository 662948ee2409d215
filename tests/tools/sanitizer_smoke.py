"""Small invocations of every kernel family, for `compute-sanitizer --tool memcheck|racecheck|initcheck`.
Sizes are tiny (the sanitizers slow kernels 10-100x); correctness is asserted against the oracle where that is cheap.
    compute-sanitizer --tool memcheck python tests/tools/sanitizer_smoke.py [part ...]      parts: movegen eval arena learner host
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mlp_ppo_2ply_multi_b200 as bg
from oracle import pyoracle as po

dev = torch.device("cuda:0")
parts = sys.argv[1:] or ["movegen", "eval", "arena", "learner", "host"]
g = np.load(os.path.join(ROOT, "tests", "golden", "values.npz"))
packed, H = g["packed"], int(g["H"])
w = bg.prepare_weights(torch.from_numpy(packed).to(dev), H)
boards, players = po.random_positions(96, seed=3)
# a few wide doubles trees so that the tail tiers run too
rng = np.random.default_rng(3)
wide = []
for k in range(6):
    b = np.zeros(52, np.int8)
    pts = rng.choice(np.arange(0, 20), size=[10, 12, 13, 15, 15, 8][k], replace=False)
    for p in pts:
        b[p] += 1
    b[int(pts[0])] += 15 - b[:24].sum()
    b[24 + 23] = 15
    wide.append(b)
boards = np.concatenate([boards, np.array(wide, np.int8)])
players = np.concatenate([players, np.zeros(6, np.uint8)])
tb, tp = torch.from_numpy(boards).to(dev), torch.from_numpy(players).to(dev)
ib, ip, ir = po.all_rolls_items(boards, players)

if "movegen" in parts:
    off, ob, om = po.movegen_batch(ib, ip, ir)
    r = bg.movegen_all_rolls(tb, tp, item_cap=4096, pool_cap=int(off[-1]) + 64)  # k_movegen21<Std>, <Big>, k_movegen<4096>
    o2, b2, _ = r.canonical()
    assert np.array_equal(o2.cpu().numpy(), off) and np.array_equal(b2.cpu().numpy(), ob)
    r = bg.movegen(torch.from_numpy(ib).to(dev), torch.from_numpy(ip).to(dev), torch.from_numpy(ir).to(dev), item_cap=4096, pool_cap=int(off[-1]) + 64,
                   want_submoves=True)  # frontier tiers 128 / 512 / 2048 / 4096 with sub-move histories
    o2, b2, m2 = r.canonical()
    assert np.array_equal(b2.cpu().numpy(), ob) and np.array_equal(m2.cpu().numpy(), om)
    r = bg.movegen(torch.from_numpy(ib).to(dev), torch.from_numpy(ip).to(dev), torch.from_numpy(ir).to(dev), item_cap=4096, pool_cap=int(off[-1]) + 64)
    o2, b2, _ = r.canonical()  # 128-node tier -> k_movegen21 single-roll tiers
    assert np.array_equal(b2.cpu().numpy(), ob)
    print("movegen ok", int(off[-1]), "afterstates", flush=True)
if "eval" in parts:
    off, ob, _ = po.movegen_batch(ib, ip, ir, want_moves=False)
    fl = np.repeat(ip, np.diff(off))
    rows = np.concatenate([ob] * (33000 // len(ob) + 1))[:33000]  # >= 32,768 rows: the tcgen05 kernel
    flg = np.concatenate([fl] * (33000 // len(ob) + 1))[:33000]
    v = bg.evaluate(torch.from_numpy(rows).to(dev), torch.from_numpy(flg).to(dev), w).cpu().numpy()
    assert bg._lib.lib().bg_eval_tc_status() == 0
    ref = po.value(packed, H, rows[:4000], flg[:4000])
    assert np.abs(v[:4000] - ref).max() < 1e-5
    v2 = bg.evaluate(torch.from_numpy(rows[:3000]).to(dev), torch.from_numpy(flg[:3000]).to(dev), w).cpu().numpy()  # CUDA-core kernel
    assert np.abs(v2 - ref[:3000]).max() < 1e-5
    f = bg.encode(torch.from_numpy(rows[:1001]).to(dev), torch.from_numpy(flg[:1001]).to(dev)).cpu().numpy()
    assert np.array_equal(f.view(np.uint32), po.encode(rows[:1001], flg[:1001]).view(np.uint32))
    a = bg.select(torch.from_numpy(v[: off[-1]]).to(dev), torch.from_numpy(off[:-1]).to(dev), torch.from_numpy(np.diff(off).astype(np.int32)).to(dev), temperature=1.0, seed=1)
    sc, nrep = bg.two_ply(torch.from_numpy(ob[:32]).to(dev), torch.from_numpy(fl[:32]).to(dev), torch.from_numpy(v[:32]).to(dev), w)
    print("eval / encode / select / two_ply ok", flush=True)
if "arena" in parts:
    ar = bg.Arena(64, hidden_size=H, device=dev, seed=5, ring_experiences=64 * 320, ring_episodes=256)
    ar.set_weights(torch.from_numpy(packed).to(dev), version=1)
    ar.reset()
    ar.step(12)
    ar.set_lookahead(4, 5, 1.0, 0.9)
    ar.step(2, lookahead=2)
    batch = ar.drain(max_episodes=64)
    print("arena ok", ar.stats()["steps"], flush=True)
    ar.close()
if "learner" in parts:
    obs, flg, rew, off, _ = po.selfplay_episodes(packed, H, 6, temperature=1.5, seed=2)
    L = bg.TD0Learner(H, dev)
    L.set_parameters(torch.from_numpy(packed).to(dev), reset_optimizer=True)
    met = L.update(torch.from_numpy(obs).to(dev), torch.from_numpy(flg).to(dev), torch.from_numpy(rew).to(dev), torch.from_numpy(off).to(dev))
    torch.cuda.synchronize()
    print("learner ok", float(met[:, 0].mean()), flush=True)
if "host" in parts:
    hb, hp = torch.from_numpy(boards).pin_memory(), torch.from_numpy(players).pin_memory()
    ha = torch.empty(21 * len(boards), dtype=torch.int32).pin_memory()
    hc = torch.empty(21 * len(boards), dtype=torch.int32).pin_memory()
    pipe = bg.HostPipeline(w, items_per_chunk=40, device=dev, all_rolls=True, item_cap=4096, rows_per_item=64)
    pipe.run(hb, hp, None, ha, hc)
    torch.cuda.synchronize()
    pipe.raise_for_status()
    off, _, _ = po.movegen_batch(ib, ip, ir, want_moves=False)
    assert np.array_equal(hc.numpy(), np.diff(off))
    pipe.close()
    print("host pipeline ok", flush=True)
print("sanitizer smoke done")
