"""GPU: sizes and inputs outside the comfortable middle -- a million envs, leading dimension > n, non-finite state,
actions outside the box, per-env parameter batches through the fused rollout."""
import numpy as np
import pytest

from helpers import REL_TOL, rel_err

pytestmark = pytest.mark.gpu


def test_one_million_envs_rollout_and_step():
    """BASELINE.json configs[3] per-GPU size and beyond (1,048,576 envs on one GPU): counters add up, results equal
    the small-batch results for the same global env ids."""
    import torch
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv
    n = 1 << 20
    big = BatchedRendezvousEnv(n, seed=5)
    big.reset()
    big.rollout(12, action_seed=9)
    big.step(torch.zeros((n, 6), dtype=torch.float32, device=big.device))
    st = big.read_stats()
    assert st["steps"] == 13 * n and st["failures"] == 0
    small = BatchedRendezvousEnv(4096, seed=5, env_offset=n - 4096)
    small.reset()
    small.rollout(12, action_seed=9)
    small.step(torch.zeros((4096, 6), dtype=torch.float32, device=big.device))
    assert torch.equal(big.get_state()[n - 4096:], small.get_state())
    assert torch.equal(big.obs[n - 4096:], small.obs)


def test_non_finite_state_is_counted_not_hung():
    """A NaN state must end the episode ('obs' reason: not in the Box), be counted as an integrator failure, and
    be replaced by a fresh episode by auto-reset -- never an endless loop."""
    import torch
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv
    for use_rollout in (False, True):
        env = BatchedRendezvousEnv(256, seed=1)
        env.reset()
        env.state_view("qc")[3] = float("nan")
        env.state_view("rc")[7] = float("inf")
        if use_rollout:
            out = env.rollout(1, actions=torch.zeros((1, 256, 6), dtype=torch.float64, device=env.device),
                              record_dones=True)
            done = out["dones"][0]
        else:
            _, _, done = env.step(torch.zeros((256, 6), dtype=torch.float64, device=env.device))
        torch.cuda.synchronize()
        assert int(done[3]) == 1 and int(done[7]) == 1
        st = env.read_stats()
        assert st["failures"] >= 1 and st["end_obs"] >= 2
        assert torch.isfinite(env.get_state()).all()            # both envs restarted from a fresh reset
        assert torch.isfinite(env.obs).all()


def test_actions_outside_the_box_are_applied_unclipped():
    """The env does not clip (rendezvous_env.py:168-173; SB3 clips before calling): |a| > 1 follows the oracle."""
    import torch
    from oracle import c_oracle as CO
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv
    n = 128
    env = BatchedRendezvousEnv(n, seed=2, auto_reset=False)
    orc = CO.COracleBatch(CO.make_params(), n)
    env.reset()
    orc.reset_from_uniforms(CO.philox_uniforms(2, np.arange(n), 1))
    a = np.random.default_rng(0).uniform(-3, 3, (n, 6))
    for _ in range(3):
        env.step(torch.as_tensor(a, device=env.device))
        orc.step(a)
    assert rel_err(env.get_state().cpu().numpy(), orc.state) <= REL_TOL


def test_param_batches_through_fused_rollout():
    """Sensitivity-style parameter batches (BASELINE.json configs[4]) stepped by rdv_rollout == separate envs."""
    import torch
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv
    batches = [(64, dict(h=400e3)), (128, dict(koz_radius=10.0, corridor_half_angle=float(np.radians(15)))),
               (64, dict(dt=0.5, t_max=30))]
    n = sum(c for c, _ in batches)
    env = BatchedRendezvousEnv(n, seed=8, param_batches=batches)
    env.reset()
    env.rollout(30, action_seed=3)
    lo = 0
    for c, kw in batches:
        e = BatchedRendezvousEnv(c, seed=8, env_offset=lo, **kw)
        e.reset()
        e.rollout(30, action_seed=3)
        assert torch.equal(env.get_state()[lo:lo + c], e.get_state())
        assert torch.equal(env.obs[lo:lo + c], e.obs)
        lo += c
    assert env.read_stats()["steps"] == 30 * n


def test_leading_dimension_larger_than_n_via_c_abi():
    """ld > n (a view into a larger allocation) through the raw C ABI."""
    import ctypes as C
    import torch
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv
    a = BatchedRendezvousEnv(100, seed=4)
    b = BatchedRendezvousEnv(100, seed=4, ld=4096)
    for e in (a, b):
        e.reset()
        e.rollout(10, action_seed=1)
        e.step(torch.full((100, 6), 0.25, dtype=torch.float64, device=e.device))
    assert b.f64.shape[1] == 4096
    assert torch.equal(a.get_state(), b.get_state()) and torch.equal(a.obs, b.obs)
