"""GPU tests of the reference-facing surfaces: the single-env Gym facade, the SB3 VecEnv, the fused policy
forward and the Monte-Carlo evaluator (golden = the reference's published workbook + its re-run here)."""
import copy

import numpy as np
import pytest

from helpers import NO_RANGE, REL_TOL, golden, rel_err

pytestmark = pytest.mark.gpu


def _policy():
    import os
    from helpers import GOLDEN
    from reinforcement_learning_rendezvous_b200 import MlpPolicy
    return MlpPolicy.load(os.path.join(GOLDEN, "policy.npz"))


# ------------------------------------------------------------------------------------------ Gym facade
def test_facade_follows_reference_episode():
    """RendezvousEnv (single-env API) replays a golden reference episode: attributes, obs, reward, done, info."""
    from reinforcement_learning_rendezvous_b200 import RendezvousEnv
    g = golden("traj_f64.npz")
    c = 2                                                # 'uniform' action case
    env = RendezvousEnv(dt=1, t_max=60, quiet=True, **NO_RANGE)
    assert env.rc is None and env.t is None
    obs = env.reset()
    assert obs.dtype == np.float32 and obs.shape == (17,)
    assert env.t == 0 and env.collided is False and env.success == 0 and env.bubble_radius == 20
    np.testing.assert_allclose(env.rc, [0, -10, 0])
    ic = g["ic"][c]
    env.rc, env.vc, env.qc, env.wc, env.qt, env.wt = ic[0:3], ic[3:6], ic[6:10], ic[10:13], ic[13:17], ic[17:20]
    np.testing.assert_array_equal(env.get_observation(), g["obs0"][c])
    for k in range(int(g["length"][c])):
        obs, rew, done, info = env.step(g["actions"][c, k])
        state = np.hstack([env.rc, env.vc, env.qc, env.wc, env.qt, env.wt])
        assert rel_err(state, g["state"][c, k]) <= REL_TOL
        assert abs(rew - g["rew"][c, k]) <= REL_TOL * max(1, abs(rew))
        assert isinstance(rew, float) and isinstance(done, bool) and done == bool(g["done"][c, k])
        assert np.abs(obs - g["obs"][c, k]).max() <= 1.2e-7
        assert set(info) == {"observation", "reward", "done", "action"}
        assert env.t == g["t"][c, k] and isinstance(env.t, int)
        assert env.bubble_radius == g["bubble"][c, k]
        assert env.collided == bool(g["collided"][c, k]) and env.success == g["success"][c, k]
        assert rel_err(env.get_errors(), g["errors"][c, k]) <= REL_TOL
        assert rel_err(env.dist_from_koz(), g["koz"][c, k]) <= REL_TOL
        assert env.check_collision() == bool(g["collision_now"][c, k])
        assert rel_err(env.total_delta_v, g["tdv"][c, k]) <= REL_TOL
    assert done


def test_facade_constants_and_helpers():
    from reinforcement_learning_rendezvous_b200 import RendezvousEnv, make_env, copy_env
    env = make_env(None, quiet=True, config=dict(rc0=30, wt0=0.02, dt=0.5, t_max=50), stochastic=False)
    np.testing.assert_allclose(env.nominal_rc0, [0, -30, 0])
    np.testing.assert_allclose(env.nominal_wt0, [0, 0, 0.02])
    assert env.max_axial_distance == 40 and env.bubble_radius0 == 40 and env.dt == 0.5 and env.t_max == 50
    assert env.max_delta_v == 0.05 and abs(env.max_delta_w - 0.006000000000000001) < 1e-18
    assert abs(env.n - 0.001039679077003123) < 1e-18
    assert env.observation_space.shape == (17,) and env.action_space.shape == (6,)
    env.reset()
    # frame transforms are rotations: R^T R v = v; target2lvlh(rd) is the goal position
    v = np.array([0.3, -1.2, 0.7])
    np.testing.assert_allclose(env.lvlh2chaser(env.chaser2lvlh(v)), v, atol=1e-14)
    np.testing.assert_allclose(env.lvlh2target(env.target2lvlh(v)), v, atol=1e-14)
    np.testing.assert_allclose(env.get_goal_pos(), env.target2lvlh(env.rd), atol=0)
    assert abs(env.get_pos_error() - np.linalg.norm(env.rc - env.get_goal_pos())) < 1e-12
    # deepcopy gives an independent env with the same future (environment_utils.copy_env)
    env.step(np.array([0.1, 0.2, -0.1, 0.3, 0.0, -0.2]))
    twin = copy_env(env)
    a = np.array([0.5, -0.5, 0.25, -0.25, 0.1, 0.0])
    o1, r1, d1, _ = env.step(a)
    assert twin.t == env.t - 0.5
    o2, r2, d2, _ = twin.step(a)
    np.testing.assert_array_equal(o1, o2)
    assert r1 == r2 and d1 == d2 and twin.t == env.t
    np.testing.assert_array_equal(twin.qc, env.qc)
    # invalid configuration: the reference's constructor assert (rendezvous_env.py:155)
    with pytest.raises(ValueError):
        RendezvousEnv(koz_radius=2)
    # in-place edits of the state arrays are honoured like on the reference's plain object
    env.rc[0] += 1.0
    assert abs(env.get_observation()[0] - (env.rc[0] / 40)) < 1e-6


def test_facade_monte_carlo_evaluate_matches_reference_rerun():
    """monte_carlo.evaluate (single-env API + predict) on 12 CSV rows vs the reference's own evaluate re-run."""
    from reinforcement_learning_rendezvous_b200 import evaluate, make_env
    mc = golden("mc.npz")
    pol = _policy()
    env = make_env(None, quiet=True, config=dict(dt=1, t_max=60), stochastic=False)
    for i in range(12):
        s = mc["ic_raw"][i]
        init = dict(rc=s[0:3].copy(), vc=s[3:6].copy(), qc=s[6:10] / np.linalg.norm(s[6:10]), wc=s[10:13].copy(),
                    qt=s[13:17] / np.linalg.norm(s[13:17]), wt=s[17:20].copy())
        out = evaluate(pol, env, init)
        for k in ("ep_len", "num_collisions", "collided", "num_successes", "succeeded"):
            assert out[k] == mc["rerun_" + k][i], (i, k)
        for k in ("total_reward", "total_delta_v", "min_dist_from_koz", "pos_error", "vel_error", "att_error",
                  "rot_error"):
            assert abs(out[k] - mc["rerun_" + k][i]) <= 2e-4 * max(1.0, abs(mc["rerun_" + k][i])), (i, k)


# ------------------------------------------------------------------------------------------ Monte Carlo
def test_device_side_monte_carlo_equals_host_loop():
    """evaluate_batch = ONE policy-fused rollout launch in evaluator mode + one [M, 16] read-back.  It must agree with
    the evaluation assembled from per-step launches and a numpy reduction of the recorded per-step arrays
    (tests/mc_hostloop.py, round 1's product path; same rule as monte_carlo.py:159-189): discrete columns equal,
    continuous columns to rounding (the fused kernel and the lane-pair step kernel order a few operations
    differently; the running sums here are sequential, numpy's are pairwise)."""
    from mc_hostloop import evaluate_batch_hostloop
    from reinforcement_learning_rendezvous_b200 import evaluate_batch, _native as N
    mc = golden("mc.npz")
    pol = _policy()
    dev_out = evaluate_batch(pol, mc["ic_raw"], return_raw=True)
    host_out = evaluate_batch_hostloop(pol, mc["ic_raw"])
    raw = dev_out.pop("raw")
    assert raw.shape == (1000, N.MC_NCOL) and (raw[:, N.MC_END_REASON] >= 0).all()
    same = np.asarray(dev_out["ep_len"]) == np.asarray(host_out["ep_len"])
    # a 1e-7 action difference between the two actor launches can move an episode end: allow 2 of 1000
    assert same.sum() >= 998, int((~same).sum())
    for k in ("num_collisions", "collided", "num_successes", "succeeded"):
        assert int((np.asarray(dev_out[k])[same] != np.asarray(host_out[k])[same]).sum()) <= 1, k
    close = same & (np.asarray(dev_out["num_successes"]) == np.asarray(host_out["num_successes"]))
    for k in ("total_reward", "total_delta_v", "min_dist_from_koz", "pos_error", "vel_error", "att_error", "rot_error"):
        a, b = np.asarray(dev_out[k])[close], np.asarray(host_out[k])[close]
        rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-2)
        assert np.quantile(rel, 0.99) <= 1e-6 and np.median(rel) <= 1e-9, (k, float(np.median(rel)), float(rel.max()))
    # the level / tail-count bookkeeping of the single running sum
    assert set(np.unique(raw[:, N.MC_LEVEL])) <= {0.0, 1.0, 2.0, 3.0, 4.0}
    assert (raw[:, N.MC_TAIL_COUNT] >= 1).all() and (raw[:, N.MC_TAIL_COUNT] <= raw[:, N.MC_EP_LEN] + 1).all()
    none = raw[:, N.MC_LEVEL] == 4
    assert (raw[none, N.MC_TAIL_COUNT] == 1).all()


def test_monte_carlo_batch_reproduces_published_workbook():
    """All 1000 published initial conditions as one GPU batch with the shipped policy: discrete columns equal
    the reference's published results (results/data_monte_carlo_results_mlp.xlsx) up to a handful of
    policy-rounding flips; success / collision totals 545 / 166 +- 3."""
    from reinforcement_learning_rendezvous_b200 import evaluate_batch
    mc = golden("mc.npz")
    out = evaluate_batch(_policy(), mc["ic_raw"])
    n = 1000
    # at most 0.5 % of the episodes (5 of 1000) may differ through a rounding flip of the fp32 policy
    counts = {}
    for src, max_flips in (("workbook_", 5), ("rerun_", 5)):
        for k in ("ep_len", "num_collisions", "collided", "num_successes", "succeeded"):
            flips = int((np.asarray(out[k], dtype=float) != mc[src + k]).sum())
            counts[src + k] = flips
    print("episodes differing from the published workbook / the reference re-run, per column:", counts)
    assert max(counts.values()) <= 5, counts
    assert abs(int(out["succeeded"].sum()) - 545) <= 3
    assert abs(int(out["collided"].sum()) - 166) <= 3
    same = np.asarray(out["ep_len"], dtype=float) == mc["rerun_ep_len"]
    for k in ("total_delta_v", "pos_error", "vel_error"):
        rel = np.abs(out[k][same] - mc["rerun_" + k][same]) / np.maximum(np.abs(mc["rerun_" + k][same]), 1e-3)
        assert np.median(rel) < 1e-5, (k, np.median(rel))
        assert np.quantile(rel, 0.99) < 1e-2, (k, np.quantile(rel, 0.99))
    assert abs(out["total_reward"].mean() - mc["rerun_total_reward"].mean()) < 0.5


# ------------------------------------------------------------------------------------------ policy
def test_policy_forward_matches_torch_fp32():
    """rdv_policy_forward (tcgen05 / TMEM, products of fp16 halves) and rdv_policy_forward_ffma (plain fp32 FMAs) against an fp64
    evaluation of the same network, at tile-ragged sizes; tolerance = fp32 accumulation error of a 64-wide MLP."""
    import torch
    pol = _policy()
    g = golden("traj_f32.npz")
    obs = np.concatenate([g["obs0"], g["obs"][:, :20].reshape(-1, 17)])
    obs = obs[np.isfinite(obs).all(axis=1)].astype(np.float32)
    w = {k: v.double().cpu() for k, v in pol.w.items()}

    def reference(x):
        x = torch.as_tensor(x).double()
        h = torch.tanh(x @ w["w0"].T + w["b0"])
        h = torch.tanh(h @ w["w1"].T + w["b1"])
        return torch.clamp(h @ w["w2"].T + w["b2"], -1, 1).numpy()

    ref = reference(obs)
    dev_obs = torch.as_tensor(obs, device=pol.device)
    act_tc = pol.forward(dev_obs).cpu().numpy()
    act_ff = pol.forward(dev_obs, ffma=True).cpu().numpy()
    assert np.abs(act_ff - ref).max() < 5e-6                    # fp32 accumulation vs an fp64 evaluation
    assert np.abs(act_tc - ref).max() < 8e-6                    # split operands: fp32-level, not fp16/TF32-level (1e-3)
    rng = np.random.default_rng(0)
    for n in (1, 31, 127, 128, 129, 300, 4096 + 77):            # partial tiles, several tiles per CTA
        x = rng.uniform(-1, 1, (n, 17)).astype(np.float32)
        a = pol.forward(torch.as_tensor(x, device=pol.device)).cpu().numpy()
        assert a.shape == (n, 6) and np.abs(a - reference(x)).max() < 8e-6, n
    # operand range of the fp16 halves: tiny observations (lo underflows to fp16 subnormals: absolute error 2^-25 per
    # term), observations far outside the Box, and beyond the fp16 range (clamped; every first-layer tanh is saturated)
    x = rng.uniform(-1, 1, (1000, 17)).astype(np.float32)
    for scale, tol in ((1e-6, 8e-6), (1e-3, 8e-6), (30.0, 2e-5), (3000.0, 2e-5)):
        a = pol.forward(torch.as_tensor(x * np.float32(scale), device=pol.device)).cpu().numpy()
        assert np.abs(a - reference(x * np.float32(scale))).max() < tol, scale
    a = pol.forward(torch.as_tensor(x * np.float32(1e7), device=pol.device)).cpu().numpy()
    assert np.isfinite(a).all() and np.abs(a).max() <= 1.0
    a1, state = pol.predict(obs[0], deterministic=True)
    assert a1.shape == (6,) and a1.dtype == np.float32 and state is None
    np.testing.assert_array_equal(a1, act_tc[0])
    # the reference's own fp32 actions (torch CPU) for the first step of each golden episode
    first = pol.predict(g["obs0"], deterministic=True)[0]
    assert np.abs(first - g["actions"][:, 0]).max() < 8e-6


# ------------------------------------------------------------------------------------------ VecEnv
def test_vec_env_contract():
    from reinforcement_learning_rendezvous_b200 import RendezvousVecEnv
    n = 512
    venv = RendezvousVecEnv(n, seed=7, t_max=15, rich_infos=True)
    assert venv.num_envs == n and venv.observation_space.shape == (17,) and venv.action_space.shape == (6,)
    obs = venv.reset()
    assert obs.shape == (n, 17) and obs.dtype == np.float32
    rng = np.random.default_rng(0)
    finished = 0
    returns = np.zeros(n)
    lengths = np.zeros(n, dtype=int)
    for k in range(40):
        a = rng.uniform(-1, 1, (n, 6)).astype(np.float32)
        venv.step_async(a)
        obs, rew, done, infos = venv.step_wait()
        assert obs.shape == (n, 17) and rew.shape == (n,) and rew.dtype == np.float32 and done.dtype == np.bool_
        assert len(infos) == n
        returns += rew
        lengths += 1
        for i in np.flatnonzero(done):
            info = infos[i]
            assert "TimeLimit.truncated" not in info
            t_obs = info["terminal_observation"]
            assert t_obs.shape == (17,) and t_obs.dtype == np.float32
            assert not np.array_equal(t_obs, obs[i])                 # returned obs is the post-reset one
            assert info["episode"]["l"] == lengths[i]
            assert abs(info["episode"]["r"] - returns[i]) < 1e-3
            assert info["end_reason"] in ("obs", "time", "bubble", "attitude")
            returns[i], lengths[i] = 0, 0
            finished += 1
        for i in np.flatnonzero(~done)[:5]:
            assert infos[i] == {}
        # post-reset observations: position near the nominal initial state, step counter zero
        steps = venv.env.step_count.cpu().numpy()
        assert (steps[done] == 0).all() and (steps[~done] == lengths[~done]).all()
    assert finished > n
    stats = venv.read_stats()
    assert stats["episodes"] == finished and stats["steps"] == 40 * n
    assert stats["end_obs"] + stats["end_time"] + stats["end_bubble"] + stats["end_attitude"] == finished
    # get_attr / env_method / set_attr
    rc = venv.get_attr("rc", indices=[0, 5])
    assert len(rc) == 2 and rc[0].shape == (3,)
    assert venv.get_attr("dt")[0] == 1.0 and venv.get_attr("koz_radius", 3) == [5.0]
    errs = venv.env_method("get_errors", indices=[1])
    assert errs[0].shape == (4,)
    venv.set_attr("rc", np.array([1.0, -2.0, 3.0]), indices=[4])
    np.testing.assert_array_equal(venv.get_attr("rc", 4)[0], [1.0, -2.0, 3.0])
    assert venv.env_is_wrapped(object) == [False] * n
    venv.close()


def test_vec_env_outputs_never_alias_later_steps():
    """SB3's collect_rollouts keeps `_last_obs` / `_last_episode_starts` across the NEXT env.step and only then adds
    them to the rollout buffer; DummyVecEnv returns copies for that reason.  The arrays a step returns must never be
    overwritten by later steps -- however long the caller keeps them -- and every env owns its info dict."""
    import sys
    from reinforcement_learning_rendezvous_b200 import RendezvousVecEnv
    n = 700
    venv = RendezvousVecEnv(n, seed=2, t_max=9)
    rng = np.random.default_rng(0)
    def hoard():
        held = []                                         # (array, its copy at the time it was returned)
        obs0 = venv.reset()
        held.append((obs0, obs0.copy()))
        for k in range(14):                               # more steps than the pinned-block pool has blocks
            obs, rew, done, infos = venv.step(rng.uniform(-1, 1, (n, 6)).astype(np.float32))
            for a in (obs, rew, done):
                held.append((a, a.copy()))
            for i in np.flatnonzero(done)[:3]:
                t = infos[i]["terminal_observation"]
                held.append((t, t.copy()))
            for a, c in held:
                np.testing.assert_array_equal(a, c)
        assert len({id(a) for a, _ in held}) == len(held)

    hoard()
    # hoarding beyond the pool's capacity degrades to real copies, never to aliasing
    assert len(venv._pool._arrays) <= venv._pool.capacity
    # per-env dicts: writing into one env's info does not show up anywhere else
    obs, rew, done, infos = venv.step(rng.uniform(-1, 1, (n, 6)).astype(np.float32))
    running = np.flatnonzero(~done)
    infos[running[0]]["custom"] = 1
    assert all("custom" not in infos[i] for i in running[1:50])
    assert len({id(infos[i]) for i in range(n)}) == n
    # a caller that drops everything lets the pool recycle its blocks
    del obs, rew, done
    before = len(venv._pool._arrays)
    for k in range(20):
        venv.step(rng.uniform(-1, 1, (n, 6)).astype(np.float32))
    assert len(venv._pool._arrays) == before


def test_vec_env_step_equals_batched_env():
    """The numpy/pinned-memory front end returns exactly what the device API computes."""
    import torch
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv, RendezvousVecEnv
    n = 300
    venv = RendezvousVecEnv(n, seed=11)
    benv = BatchedRendezvousEnv(n, seed=11)
    np.testing.assert_array_equal(venv.reset(), benv.reset().cpu().numpy())
    rng = np.random.default_rng(3)
    for _ in range(30):
        a = rng.uniform(-1, 1, (n, 6)).astype(np.float32)
        obs, rew, done, _ = venv.step(a)
        o2, r2, d2 = benv.step(torch.as_tensor(a, device=benv.device))
        np.testing.assert_array_equal(obs, o2.cpu().numpy())
        np.testing.assert_array_equal(rew, r2.float().cpu().numpy())
        np.testing.assert_array_equal(done, d2.bool().cpu().numpy())


def test_param_batches_match_separate_envs():
    """Per-env parameter batches (sensitivity-sweep axes) == separate envs with those constructor arguments."""
    import torch
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv
    batches = [(64, dict(h=400e3)), (64, dict(koz_radius=10.0)), (96, dict(dt=0.5, corridor_half_angle=0.5)),
               (32, dict(rc0=np.array([0., -30., 0.])))]
    n = sum(c for c, _ in batches)
    env = BatchedRendezvousEnv(n, seed=5, param_batches=batches, auto_reset=True)
    env.reset()
    parts, lo = [], 0
    for c, kw in batches:
        e = BatchedRendezvousEnv(c, seed=5, env_offset=lo, auto_reset=True, **kw)
        e.reset()
        parts.append(e)
        lo += c
    rng = np.random.default_rng(1)
    for _ in range(25):
        a = torch.as_tensor(rng.uniform(-1, 1, (n, 6)), device=env.device)
        obs, rew, done = env.step(a)
        lo = 0
        for (c, _), e in zip(batches, parts):
            o2, r2, d2 = e.step(a[lo:lo + c].contiguous())
            assert torch.equal(obs[lo:lo + c], o2) and torch.equal(rew[lo:lo + c], r2)
            assert torch.equal(done[lo:lo + c], d2)
            lo += c
    st = env.read_stats()
    assert st["steps"] == 25 * n


def test_checkpoint_roundtrip():
    import torch
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv
    env = BatchedRendezvousEnv(128, seed=21)
    env.reset()
    a = torch.rand((128, 6), dtype=torch.float64, device=env.device) * 2 - 1
    for _ in range(5):
        env.step(a)
    sd = copy.deepcopy(env.state_dict())
    ref = [env.step(a)[0].clone() for _ in range(10)]
    other = BatchedRendezvousEnv(128, seed=0)
    other.load_state_dict(sd)
    for k in range(10):
        assert torch.equal(other.step(a)[0], ref[k])


def test_vec_env_step_arrays_matches_step_infos():
    from reinforcement_learning_rendezvous_b200 import RendezvousVecEnv
    n = 400
    a_env = RendezvousVecEnv(n, seed=13, t_max=12, rich_infos=True)
    b_env = RendezvousVecEnv(n, seed=13, t_max=12)
    a_env.reset(); b_env.reset()
    rng = np.random.default_rng(1)
    seen = 0
    for _ in range(30):
        act = rng.uniform(-1, 1, (n, 6)).astype(np.float32)
        obs, rew, done, infos = a_env.step(act)
        o2, r2, d2, fin = b_env.step_arrays(act)
        np.testing.assert_array_equal(obs, o2); np.testing.assert_array_equal(rew, r2); np.testing.assert_array_equal(done, d2)
        np.testing.assert_array_equal(fin["index"], np.flatnonzero(done))
        for j, i in enumerate(fin["index"]):
            info = infos[i]
            np.testing.assert_array_equal(info["terminal_observation"], fin["terminal_observation"][j])
            assert info["episode"]["l"] == fin["episode_length"][j]
            assert abs(info["episode"]["r"] - fin["episode_return"][j]) < 1e-6
            assert info["is_success"] == bool(fin["is_success"][j]) and info["collided"] == bool(fin["collided"][j])
            assert info["end_reason"] == ("obs", "time", "bubble", "attitude")[fin["end_reason"][j]]
            seen += 1
    assert seen > n


def test_sensitivity_sweep_as_one_batch():
    """BASELINE.json configs[4]: parameter sets of sensitivity_analysis.py:97-134 evaluated as blocks of one batch;
    the nominal block reproduces the policy's Monte-Carlo success rate, harder settings do worse."""
    from reinforcement_learning_rendezvous_b200 import evaluate_sweep
    pol = _policy()
    sets = [dict(), dict(h=400e3), dict(wt0=float(np.radians(2.5))), dict(koz_radius=10.0),
            dict(corridor_half_angle=float(np.radians(15))), dict(rc0=30.0)]
    res = evaluate_sweep(pol, sets, episodes_per_set=1024, seed=3, dt=1, t_max=60)
    assert len(res) == len(sets) and all(r["episodes"] == 1024 for r in res)
    nominal = res[0]
    assert 0.40 < nominal["success_rate"] < 0.70          # the published experiment: 545 / 1000
    assert 0.08 < nominal["collision_rate"] < 0.30        # 166 / 1000
    assert abs(res[1]["success_rate"] - nominal["success_rate"]) < 0.15      # altitude barely matters
    assert res[4]["collision_rate"] > nominal["collision_rate"]               # a narrower corridor collides more
    assert res[5]["success_rate"] <= nominal["success_rate"]      # 30 m out: outside what the policy was trained on
    for r in res:
        assert np.isfinite(r["mean_return"]) and 0 < r["mean_length_s"] <= 60


def test_full_sensitivity_grid_is_one_launch_and_equals_separate_envs():
    """BASELINE.json configs[4]: every valid combination of the grid of sensitivity_analysis.py:97-134 (3,750 parameter
    sets) as ONE batch -- a device table of RdvParams + one entry index per block of 32 envs, so that reset, step and
    rollout are one launch each -- gives the same bits as separate envs built with those constructor arguments, and
    agrees with the C oracle."""
    import torch
    from oracle import c_oracle as CO
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv, sensitivity_grid
    from reinforcement_learning_rendezvous_b200.environment_utils import config_to_kwargs
    grid = sensitivity_grid()
    assert len(grid) == 3750 and len(sensitivity_grid(include_invalid=True)) == 5625
    per = 32
    kws = [{k: v for k, v in config_to_kwargs(g, stochastic=True).items() if v is not None} for g in grid]
    n = per * len(grid)
    env = BatchedRendezvousEnv(n, seed=6, auto_reset=True, param_batches=[(per, kw) for kw in kws])
    assert env.param_table is not None and len(env._units) == 1          # one launch for the whole batch
    env.reset()
    g = torch.Generator(device=env.device)
    g.manual_seed(0)
    K = 6
    acts = torch.rand((K, n, 6), dtype=torch.float64, device=env.device, generator=g) * 2 - 1
    rews = []
    for k in range(K):
        rews.append(env.step(acts[k])[1].clone())
    out = env.rollout(K, actions=acts, record_rewards=True, record_dones=True)      # records with a table: allowed now
    rews2 = out["rewards"]
    state = env.get_state()
    rng = np.random.default_rng(1)
    for gi in [0, 1, len(grid) - 1] + rng.choice(len(grid), 9, replace=False).tolist():
        lo = gi * per
        e = BatchedRendezvousEnv(per, seed=6, env_offset=lo, auto_reset=True, **kws[gi])
        e.reset()
        for k in range(K):
            assert torch.equal(e.step(acts[k, lo:lo + per].contiguous())[1], rews[k][lo:lo + per]), (gi, k)
        o2 = e.rollout(K, actions=acts[:, lo:lo + per].contiguous(), record_rewards=True)
        assert torch.equal(o2["rewards"], rews2[:, lo:lo + per]), gi
        assert torch.equal(e.get_state(), state[lo:lo + per]), gi
        assert torch.equal(e.i32[:, :per], env.i32[:, lo:lo + per]), gi
        # and against the C oracle (first step from the same reset draws)
        orc = CO.COracleBatch(CO.make_params(**kws[gi]), per)
        orc.reset_from_uniforms(CO.philox_uniforms(6, lo + np.arange(per), np.ones(per, dtype=np.int32)))
        _, o_rew, _ = orc.step(acts[0, lo:lo + per].cpu().numpy())
        assert rel_err(rews[0][lo:lo + per].cpu().numpy(), o_rew) <= REL_TOL, gi
    st = env.read_stats()
    assert st["steps"] == 2 * K * n and st["failures"] == 0


def test_sweep_sharding_is_invariant():
    """evaluate_sweep over parameter sets sharded as two ranks == the unsharded sweep (global env ids key the reset
    streams), and its evaluator-mode launch agrees with a per-step evaluation of the same blocks."""
    from reinforcement_learning_rendezvous_b200 import evaluate_sweep
    pol = _policy()
    sets = [dict(), dict(h=400e3), dict(dt=0.5), dict(koz_radius=10.0), dict(rc0=20.0), dict(corridor_half_angle=0.6)]
    whole = evaluate_sweep(pol, sets, episodes_per_set=256, seed=3, t_max=60)
    parts = evaluate_sweep(pol, sets, episodes_per_set=256, seed=3, t_max=60, rank=0, world_size=2) + \
        evaluate_sweep(pol, sets, episodes_per_set=256, seed=3, t_max=60, rank=1, world_size=2)
    assert len(parts) == len(whole) == len(sets)
    for a, b in zip(whole, parts):
        assert a == b


@pytest.mark.parametrize("n", [1, 33])
def test_vec_env_tiny_batches(n):
    """The one-block host path at sizes where the finished-row bound exceeds the batch."""
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv, RendezvousVecEnv
    import torch
    venv = RendezvousVecEnv(n, seed=5, t_max=4)
    benv = BatchedRendezvousEnv(n, seed=5, t_max=4)
    np.testing.assert_array_equal(venv.reset(), benv.reset().cpu().numpy())
    rng = np.random.default_rng(0)
    ends = 0
    for _ in range(12):
        a = rng.uniform(-1, 1, (n, 6)).astype(np.float32)
        obs, rew, done, infos = venv.step(a)
        o2, r2, d2 = benv.step(torch.as_tensor(a, device=benv.device))
        np.testing.assert_array_equal(obs, o2.cpu().numpy())
        np.testing.assert_array_equal(done, d2.bool().cpu().numpy())
        for i in np.flatnonzero(done):
            np.testing.assert_array_equal(infos[i]["terminal_observation"], benv.terminal_obs[i].cpu().numpy())
            assert infos[i]["episode"]["l"] <= 4
            ends += 1
    assert ends >= 2 * n
