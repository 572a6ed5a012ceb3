"""A few policy-fused rollout launches (target for ncu; development tool)."""
import sys, torch
sys.path.insert(0, '.')
from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv, MlpPolicy
pol = MlpPolicy.load("tests/golden/policy.npz")
env = BatchedRendezvousEnv(65536, seed=0)
env.reset()
for r in range(4):
    env.rollout(16, policy=pol)
torch.cuda.synchronize()
print("ok", env.read_stats()["steps"])
