"""Small run through every kernel (target for compute-sanitizer; development tool)."""
import os, sys, numpy as np, torch
sys.path.insert(0, '.')
from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv, MlpPolicy, RendezvousEnv
for n in (1, 33, 1000):
    env = BatchedRendezvousEnv(n, seed=1, t_max=8)
    env.reset()
    for k in range(12):
        env.step(torch.rand((n, 6), dtype=torch.float64, device='cuda') * 2 - 1)
        env.step(torch.rand((n, 6), dtype=torch.float32, device='cuda') * 2 - 1)
    env.errors(); env.observe(); env.refresh_flags()
    env.reset(mask=torch.ones(n, dtype=torch.uint8))
    from reinforcement_learning_rendezvous_b200 import _native as N
    for tpb in (256, 448):
        N.lib().rdv_tune(N.TUNE_ROLLOUT_TPB, tpb)
        env.rollout(10, action_seed=3, record_rewards=True, record_dones=True, record_obs=True, record_actions=True)
        env.rollout(5, actions=torch.rand((5, n, 6), dtype=torch.float32, device='cuda'))
    N.lib().rdv_tune(N.TUNE_ROLLOUT_TPB, 0)
    aniso = BatchedRendezvousEnv(n, seed=1, inertia=np.diag([10.0, 16.0, 22.0]), chaser_torque=[1e-3, 0, 0])
    aniso.reset(); aniso.step(torch.zeros((n, 6), dtype=torch.float64, device='cuda')); aniso.rollout(3, action_seed=1)
    cf = BatchedRendezvousEnv(n, seed=1, integrator="closed_form")
    cf.reset(); cf.step(torch.zeros((n, 6), dtype=torch.float64, device='cuda')); cf.rollout(3, action_seed=1)
pol = MlpPolicy.load("tests/golden/policy.npz")
pol.forward(env.obs)
e = RendezvousEnv(quiet=True); e.reset(); e.step(np.zeros(6)); e.get_errors(); e.chaser2lvlh(np.ones(3))
torch.cuda.synchronize()
print("sanitize target ok")
