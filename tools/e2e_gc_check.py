"""Does the cyclic GC matter for RendezvousVecEnv.step at 65,536 envs? (development tool)"""
import gc, sys, time, numpy as np
sys.path.insert(0, '.')
from reinforcement_learning_rendezvous_b200 import RendezvousVecEnv
n = 65536
v = RendezvousVecEnv(n, seed=0)
v.reset()
acts = np.random.default_rng(0).uniform(-1, 1, (64, n, 6)).astype(np.float32)
for k in range(20): v.step(acts[k % 64])
def run(tag, steps=100):
    t0 = time.perf_counter()
    for k in range(steps):
        obs, rew, done, infos = v.step(acts[k % 64])
    dt = time.perf_counter() - t0
    print(f"{tag}: {1e3 * dt / steps:.2f} ms/step  {n * steps / dt / 1e6:.1f} M env-steps/s", flush=True)
run("gc enabled")
gc.disable(); run("gc disabled"); gc.enable()
gc.freeze(); run("gc enabled after freeze"); gc.unfreeze()
ballast = [{"x": i} for i in range(2_000_000)]      # a process with many live containers
run("gc enabled, 2M live dicts")
gc.disable(); run("gc disabled, 2M live dicts"); gc.enable()
