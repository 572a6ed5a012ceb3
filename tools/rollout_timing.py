"""Throughput of the fused rollout kernel vs steps-per-launch (development tool)."""
import sys, torch
sys.path.insert(0, '.')
from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
for integ in ("rk45", "closed_form"):
    env = BatchedRendezvousEnv(n, seed=0, integrator=integ)
    env.reset()
    for K in (1, 4, 16, 64, 256):
        env.rollout(K, action_seed=1)
        torch.cuda.synchronize()
        reps = max(1, 512 // K)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(reps):
            env.rollout(K, action_seed=1, step_base=1000 + r * K)
        e1.record(); torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / (reps * K)
        print(f"{integ:12s} n={n} K={K:4d}: {us:7.2f} us/step  {n / us / 1e3:7.3f} G env-steps/s", flush=True)
    st = env.read_stats()
    print("   mean episode length", st["length_sum"] / max(st["episodes"], 1), "rk/solve", st["rk_accepted"] / max(st["steps"], 1) / 2)
