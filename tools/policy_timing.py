"""Throughput of the policy-fused rollout (development tool)."""
import sys, torch
sys.path.insert(0, '.')
from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv, MlpPolicy
pol = MlpPolicy.load("tests/golden/policy.npz")
for n in (65536, 131072):
    env = BatchedRendezvousEnv(n, seed=0)
    env.reset()
    for mode in ("philox", "policy"):
        kw = dict(action_seed=1) if mode == "philox" else dict(policy=pol)
        for K in (16, 64):
            env.rollout(K, **kw); torch.cuda.synchronize()
            reps = max(1, 256 // K)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for r in range(reps): env.rollout(K, **kw)
            e1.record(); torch.cuda.synchronize()
            us = 1e3 * e0.elapsed_time(e1) / (reps * K)
            print(f"n={n} {mode:7s} K={K:3d}: {us:7.2f} us/step  {n / us / 1e3:6.3f} G env-steps/s", flush=True)
    st = env.read_stats()
    print("  episodes", st["episodes"], "mean length", st["length_sum"] / max(st["episodes"], 1), "success rate", st["succeeded"] / max(st["episodes"], 1))
# separate policy kernel for reference
obs = env.obs
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
pol.forward(obs); e0.record()
for _ in range(20): pol.forward(obs)
e1.record(); torch.cuda.synchronize()
print(f"stand-alone tcgen05 policy kernel, n={obs.shape[0]}: {1e3 * e0.elapsed_time(e1) / 20:.1f} us")
