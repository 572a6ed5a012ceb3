"""Headless ports of verification/*.py run against the CUDA kernels, with the pass criteria of SURVEY.md
section 4 and the reference's own end states (tests/golden/verify.npz) as goldens."""
import numpy as np
import pytest

from helpers import REL_TOL, golden, rel_err

pytestmark = pytest.mark.gpu


def test_verify_cw():
    from reinforcement_learning_rendezvous_b200 import verification as V
    out, g = V.verify_cw(), golden("verify.npz")
    assert out["max_pos_diff"] < 1e-11 and out["max_vel_diff"] < 1e-13       # stepping vs one-shot analytic CW
    assert rel_err(out["r"], g["cw_r"]) <= REL_TOL and rel_err(out["v"], g["cw_v"], floor=1e-2) <= REL_TOL
    np.testing.assert_allclose(out["r_analytic"], g["cw_r_analytic"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(out["r"][-1], [-1.16989553, 0.17113477, 0.70741817], atol=1e-8)


def test_verify_cw2():
    from reinforcement_learning_rendezvous_b200 import verification as V
    out, g = V.verify_cw2(), golden("verify.npz")
    assert out["action"][0] == g["cw2_action"][0]
    assert out["max_vy_dev"] < 2e-4
    assert rel_err(out["state"], g["cw2_state"]) <= REL_TOL


def test_verify_attitude_torque():
    from reinforcement_learning_rendezvous_b200 import verification as V
    out, g = V.verify_attitude_torque(), golden("verify.npz")
    assert rel_err(out["state"], g["torque_state"]) <= REL_TOL
    assert abs(out["w_final"][2] - out["w_ideal"]) < 1e-12 and abs(np.degrees(out["w_final"][2]) - 11.172677) < 1e-6
    # impulsive-at-step-start model: theta - ideal = 0.5 * w_final * dt
    assert abs((out["theta"] - out["theta_ideal"]) - 0.5 * out["w_final"][2] * out["dt"]) < 1e-9
    assert abs(np.degrees(out["theta"]) - 184.349171) < 1e-5


def test_verify_attitude_racket():
    from reinforcement_learning_rendezvous_b200 import verification as V
    out, g = V.verify_attitude_racket(), golden("verify.npz")
    assert out["wc_drift"] < 1e-15                                  # isotropic inertia: no flip possible
    assert rel_err(out["state"], g["racket_state"]) <= REL_TOL
    np.testing.assert_allclose(out["state"][-1, 6:10], [-0.173713492, 0, 0.984794264, 0.00196958853], atol=1e-8)


def test_verify_attitude_general_rigid_body():
    """General diagonal inertia (+ held torque): RK45 replica (rtol 1e-7, atol 1e-6) vs a fine RK4."""
    from reinforcement_learning_rendezvous_b200 import verification as V
    out = V.verify_attitude(steps=120)
    assert out["max_q_diff"] < 2e-5 and out["max_w_diff"] < 2e-6, (out["max_q_diff"], out["max_w_diff"])
    out = V.verify_attitude(steps=60, torque=np.array([0.002, -0.001, 0.0015]))
    assert out["max_q_diff"] < 2e-5 and out["max_w_diff"] < 2e-6, (out["max_q_diff"], out["max_w_diff"])
    # intermediate-axis rotation with a non-isotropic body does tumble (what verify_attitude_racket was after)
    out = V.verify_attitude(steps=400, inertia=np.diag([10.0, 16.0, 22.0]), wc0=np.radians([0.05, 5.0, 0.05]))
    assert np.abs(out["state"][:, 5]).min() < 0.5 * np.abs(out["state"][0, 5])
    assert out["max_w_diff"] < 2e-5


def test_general_inertia_matches_c_oracle():
    """Non-isotropic bodies + held torque through the batched API vs the C oracle (same RK45 restatement)."""
    import torch
    from oracle import c_oracle as CO
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv
    kw = dict(inertia=np.diag([10.0, 16.0, 22.0]), inertia_target=np.array([[20, 1, 0], [1, 15, 2], [0, 2, 12.0]]),
              chaser_torque=np.array([0.002, -0.001, 0.0015]))
    n = 512
    env = BatchedRendezvousEnv(n, seed=4, auto_reset=False, **kw)
    orc = CO.COracleBatch(CO.make_params(**kw), n)
    env.reset()
    orc.reset_from_uniforms(CO.philox_uniforms(4, np.arange(n), 1))
    rng = np.random.default_rng(2)
    for _ in range(20):
        a = 0.5 * rng.uniform(-1, 1, (n, 6))
        _, rew, done = env.step(torch.as_tensor(a, device=env.device))
        _, o_rew, o_done = orc.step(a, threads=8)
        assert rel_err(env.get_state().cpu().numpy(), orc.state) <= REL_TOL
        assert rel_err(rew.cpu().numpy(), o_rew) <= REL_TOL
        np.testing.assert_array_equal(done.cpu().numpy(), o_done)


def test_initial_state_distribution_report():
    from reinforcement_learning_rendezvous_b200 import verification as V
    rep = V.initial_state_distribution(100_000, seed=0)
    for k, r in rep.items():
        assert r["max"] <= r["range"] * (1 + 1e-9)
        assert abs(r["mean"] / r["expected_mean"] - 1) < 0.02, k
        assert abs(r["var"] / r["expected_var"] - 1) < 0.05, k
