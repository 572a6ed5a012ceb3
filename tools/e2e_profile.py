"""Where the host time of RendezvousVecEnv.step goes (development tool)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from reinforcement_learning_rendezvous_b200 import RendezvousVecEnv
n = 65536
for rich in (True, False):
    v = RendezvousVecEnv(n, seed=0, rich_infos=rich)
    v.reset()
    acts = np.random.default_rng(0).uniform(-1, 1, (8, n, 6)).astype(np.float32)
    for k in range(20): v.step(acts[k % 8])
    t0 = time.perf_counter()
    ta = tw = 0.0
    for k in range(100):
        a0 = time.perf_counter(); v.step_async(acts[k % 8]); a1 = time.perf_counter(); v.step_wait(); a2 = time.perf_counter()
        ta += a1 - a0; tw += a2 - a1
    dt = time.perf_counter() - t0
    print(f"rich_infos={rich}: {1e3*dt/100:.2f} ms/step  (async {1e3*ta/100:.2f}, wait {1e3*tw/100:.2f})  {n*100/dt/1e6:.1f} M env-steps/s")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for k in range(30): v.step(acts[k % 8])
pr.disable(); pstats.Stats(pr).sort_stats("cumtime").print_stats(12)
