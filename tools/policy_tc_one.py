"""A few launches of the tcgen05 policy kernel at 131,072 envs (target for ncu; development tool)."""
import sys, torch
sys.path.insert(0, '.')
from reinforcement_learning_rendezvous_b200 import MlpPolicy
pol = MlpPolicy.load("tests/golden/policy.npz")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
g = torch.Generator(device='cuda'); g.manual_seed(0)
obs = torch.rand((n, 17), dtype=torch.float32, device='cuda', generator=g) * 2 - 1
for _ in range(4):
    a = pol.forward(obs)
torch.cuda.synchronize()
print("ok", float(a.abs().sum()))
