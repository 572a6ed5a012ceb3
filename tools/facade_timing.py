"""Steps per second of the single-env Gym facade (RendezvousEnv: reset()/step(a) -> numpy obs, float, bool, dict), the
drop-in for the reference's DummyVecEnv-of-one usage (main.py:33-34, monte_carlo.py:94-207).  Development tool."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from reinforcement_learning_rendezvous_b200 import RendezvousEnv
env = RendezvousEnv(quiet=True)
env.reset()
rng = np.random.default_rng(0)
acts = rng.uniform(-1, 1, (4096, 6))
for k in range(200):
    _, _, done, _ = env.step(acts[k])
    if done:
        env.reset()
t0 = time.perf_counter()
S = 3000
for k in range(S):
    _, _, done, _ = env.step(acts[k % 4096])
    if done:
        env.reset()
dt = time.perf_counter() - t0
print(f"RendezvousEnv facade: {S / dt:.0f} env-steps/s ({1e6 * dt / S:.0f} us per step incl. resets)")
