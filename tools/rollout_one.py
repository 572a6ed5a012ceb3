"""A few fused-rollout launches at a given size (target for ncu captures; development tool)."""
import sys, torch
sys.path.insert(0, '.')
from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv
n = int(sys.argv[1]); K = int(sys.argv[2]) if len(sys.argv) > 2 else 16
env = BatchedRendezvousEnv(n, seed=0)
env.reset()
for r in range(6):
    env.rollout(K, action_seed=1, step_base=r * K)
torch.cuda.synchronize()
print("ok", env.read_stats()["steps"])
