"""GPU parity of rdv_reset: the uniform-draw -> initial-state map against the reference (golden reset.npz),
the Philox4x32-10 stream against the C oracle, and the distribution checks of
verification/initial_state_distribution.py:88-123 (mean = range/2, var = range^2/12 of the six deviation
magnitudes; cube-normalised directions)."""
import numpy as np
import pytest

from helpers import REL_TOL, cfg_kwargs, golden, rel_err

pytestmark = pytest.mark.gpu


def test_reset_from_reference_uniforms():
    import torch
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv
    g = golden("reset.npz")
    cfgs = [str(c) for c in g["cfg"]]
    for cfg in sorted(set(cfgs)):
        idx = np.array([i for i, c in enumerate(cfgs) if c == cfg])
        kw, reward = cfg_kwargs(cfg)
        env = BatchedRendezvousEnv(len(idx), reward_kwargs=reward, **kw)
        obs = env.reset(uniforms=torch.as_tensor(g["uniforms"][idx]))
        assert rel_err(env.get_state().cpu().numpy(), g["state"][idx]) <= REL_TOL
        assert np.abs(obs.cpu().numpy() - g["obs"][idx]).max() <= 1.2e-7
        np.testing.assert_array_equal(env.collided.cpu().numpy(), g["collided"][idx])
        np.testing.assert_array_equal(env.success.cpu().numpy(), g["success"][idx])
        assert (env.step_count.cpu().numpy() == 0).all()
        assert (env.total_delta_v.cpu().numpy() == 0).all()


def test_masked_reset_leaves_other_envs_untouched():
    import torch
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv
    env = BatchedRendezvousEnv(256, seed=9)
    env.reset()
    env.step(torch.zeros((256, 6), dtype=torch.float64, device=env.device))
    before = env.get_state().clone()
    ep_before = env.episode_index.clone()
    mask = torch.zeros(256, dtype=torch.uint8)
    mask[::3] = 1
    env.reset(mask=mask)
    after = env.get_state()
    keep = ~mask.bool().to(env.device)
    assert torch.equal(after[keep], before[keep])
    assert not torch.equal(after[~keep], before[~keep])
    assert torch.equal(env.episode_index[keep], ep_before[keep])
    assert torch.equal(env.episode_index[~keep], ep_before[~keep] + 1)
    assert (env.step_count[~keep] == 0).all() and (env.step_count[keep] == 1).all()


def test_philox_stream_matches_c_oracle_and_is_shard_invariant():
    """(seed, global env id, episode index) alone determines a reset: two shards [0,512) + [512,1024) give
    the same states as one 1024-env batch, and both equal the C oracle's Philox."""
    from oracle import c_oracle as CO
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv
    seed = 0xDEADBEEF12345
    whole = BatchedRendezvousEnv(1024, seed=seed)
    whole.reset()
    a = BatchedRendezvousEnv(512, seed=seed, env_offset=0)
    b = BatchedRendezvousEnv(512, seed=seed, env_offset=512)
    a.reset()
    b.reset()
    s = whole.get_state().cpu().numpy()
    np.testing.assert_array_equal(s[:512], a.get_state().cpu().numpy())
    np.testing.assert_array_equal(s[512:], b.get_state().cpu().numpy())
    orc = CO.COracleBatch(CO.make_params(), 1024)
    orc.reset_from_uniforms(CO.philox_uniforms(seed, np.arange(1024), 1))
    assert rel_err(s, orc.state) <= REL_TOL
    # Philox4x32-10 known-answer vectors (Random123 kat_vectors): counter/key all zero and all ones
    np.testing.assert_array_equal(CO.philox_raw([0, 0, 0, 0], 0, 0),
                                  np.array([0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8], dtype=np.uint32))
    np.testing.assert_array_equal(CO.philox_raw([0xffffffff] * 4, 0xffffffff, 0xffffffff),
                                  np.array([0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd], dtype=np.uint32))


def test_initial_state_distribution():
    """verification/initial_state_distribution.py:88-123 on 262,144 GPU resets."""
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv
    n = 1 << 18
    env = BatchedRendezvousEnv(n, seed=2024)
    env.reset()
    s = env.get_state().cpu().numpy()
    p = env.params
    rc_dev = s[:, 0:3] - np.array(p.rc0[:])
    vc_dev = s[:, 3:6] - np.array(p.vc0[:])
    qc, wc, qt, wt = s[:, 6:10], s[:, 10:13], s[:, 13:17], s[:, 17:20]
    mags = {
        "rc": (np.linalg.norm(rc_dev, axis=1), p.rc0_range),
        "vc": (np.linalg.norm(vc_dev, axis=1), p.vc0_range),
        "qc": (2 * np.arccos(np.clip(np.abs(qc[:, 0]), 0, 1)), p.qc0_range),
        "wc": (np.linalg.norm(wc, axis=1), p.wc0_range),
        "qt": (2 * np.arccos(np.clip(np.abs(qt[:, 0]), 0, 1)), p.qt0_range),
        "wt": (np.linalg.norm(wt, axis=1), p.wt0_range),
    }
    for name, (m, rng) in mags.items():
        sigma = rng / np.sqrt(12)
        assert m.max() <= rng * (1 + 1e-9) and m.min() >= 0, name
        assert abs(m.mean() - rng / 2) < 4 * sigma / np.sqrt(n), name
        assert abs(m.var() / (rng ** 2 / 12) - 1) < 0.02, name
    assert np.abs(np.linalg.norm(qc, axis=1) - 1).max() < 1e-14
    assert np.abs(np.linalg.norm(qt, axis=1) - 1).max() < 1e-14
    # cube-normalised directions (utils/general.py:248-254): P(|u_x| > 1/sqrt(3)) is higher than for a
    # sphere-uniform direction; compare against a numpy sample of the same construction
    d = rc_dev / np.linalg.norm(rc_dev, axis=1, keepdims=True)
    ref = np.random.default_rng(0).uniform(-1, 1, (n, 3))
    ref /= np.linalg.norm(ref, axis=1, keepdims=True)
    for axis in range(3):
        h_gpu, _ = np.histogram(d[:, axis], bins=20, range=(-1, 1))
        h_ref, _ = np.histogram(ref[:, axis], bins=20, range=(-1, 1))
        assert np.abs(h_gpu - h_ref).max() < 6 * np.sqrt(h_ref.max()), axis
    # successive episodes of one env are independent draws
    first = s[:, 0].copy()
    env.reset()
    second = env.get_state().cpu().numpy()[:, 0]
    assert abs(np.corrcoef(first, second)[0, 1]) < 0.01
