"""Where the host time of RendezvousVecEnv.step goes (development tool): per-phase wall-clock of 300 steps."""
import sys, time, gc
import numpy as np, torch
sys.path.insert(0, '.')
from reinforcement_learning_rendezvous_b200 import RendezvousVecEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
v = RendezvousVecEnv(n, seed=0)
v.reset()
rng = np.random.default_rng(0)
ring = rng.uniform(-1, 1, (16, n, 6)).astype(np.float32)
for k in range(30):
    v.step(ring[k % 16])
T = dict(stage=0.0, launch_fetch=0.0, infos=0.0, total=0.0)
S = 300
worst = 0.0
for k in range(S):
    t0 = time.perf_counter()
    v.step_async(ring[k % 16])
    t1 = time.perf_counter()
    gc.disable()
    obs, rew, done, rows = v._launch_and_fetch(renew_infos=True)   # last step's info dicts renewed while the GPU works
    t2 = time.perf_counter()
    infos = v._build_infos(rows)
    gc.enable()
    del rows
    t3 = time.perf_counter()
    T["stage"] += t1 - t0; T["launch_fetch"] += t2 - t1; T["infos"] += t3 - t2; T["total"] += t3 - t0
    worst = max(worst, t3 - t0)
print({k: round(1e3 * x / S, 4) for k, x in T.items()}, "ms per step; worst step", round(1e3 * worst, 3), "ms; finished/step",
      int(done.sum()), "pool blocks", len(v._pool._arrays), "extra fetches", v.extra_fetches)
t0 = time.perf_counter()
for k in range(S):
    v.step(ring[k % 16])
print("v.step as a whole: %.4f ms per step" % (1e3 * (time.perf_counter() - t0) / S))
# the pieces of launch_fetch
env = v.env
torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(S):
    env.step(v._d_act32)
torch.cuda.synchronize()
print("rdv_step launch + execute: %.4f ms per step" % (1e3 * (time.perf_counter() - t0) / S))
t, base, _ = v._pool.acquire()
L = v.layout
nb = L["rows"] + 128 * 4096
t0 = time.perf_counter()
for k in range(S):
    t[:nb].copy_(env.host_block[:nb], non_blocking=True)
    torch.cuda.current_stream().synchronize()
print("D2H of %.2f MB + sync: %.4f ms" % (nb / 1e6, 1e3 * (time.perf_counter() - t0) / S))
t0 = time.perf_counter()
for k in range(S):
    v._d_act32.copy_(v._h_act32, non_blocking=True)
torch.cuda.synchronize()
print("H2D actions: %.4f ms" % (1e3 * (time.perf_counter() - t0) / S))
