"""GPU parity of rdv_rollout (K fused steps per launch): against the per-step API (bit-identical), against the
C oracle with the shared Philox action + reset streams, and the per-step records."""
import numpy as np
import pytest

from helpers import REL_TOL, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_rollout_equals_repeated_step(dtype):
    """Same actions through rollout() and through K step() calls: identical state, records and statistics."""
    import torch
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv
    n, K = 1000, 48
    a = BatchedRendezvousEnv(n, seed=3, env_offset=77, t_max=25)
    b = BatchedRendezvousEnv(n, seed=3, env_offset=77, t_max=25)
    a.reset()
    b.reset()
    g = torch.Generator(device=a.device)
    g.manual_seed(0)
    acts = (torch.rand((K, n, 6), dtype=torch.float64, device=a.device, generator=g) * 2 - 1).to(getattr(torch, dtype))
    out = a.rollout(K, actions=acts, record_rewards=True, record_dones=True, record_obs=True)
    for k in range(K):
        obs, rew, done = b.step(acts[k])
        # the thread-per-env rollout and the lane-pair step kernel order a few operations differently
        assert rel_err(out["rewards"][k].cpu().numpy(), rew.cpu().numpy()) <= 1e-12
        assert torch.equal(out["dones"][k], done)
        assert (out["obs_steps"][k] - obs).abs().max().item() <= 1.2e-7
    assert rel_err(a.get_state().cpu().numpy(), b.get_state().cpu().numpy()) <= 1e-11
    assert torch.equal(a.i32[:, :n], b.i32[:, :n])
    assert (out["obs"] - b.obs).abs().max().item() <= 1.2e-7
    sa, sb = a.read_stats(), b.read_stats()
    for key in ("steps", "episodes", "succeeded", "collided", "end_obs", "end_time", "end_bubble", "end_attitude",
                "rk_accepted", "rk_rejected", "failures"):
        assert sa[key] == sb[key], key
    assert sa["episodes"] > n and abs(sa["reward_sum"] - sb["reward_sum"]) <= 1e-9 * abs(sb["reward_sum"])


@pytest.mark.parametrize("cfg", [dict(dt=0.25, t_max=10), dict(dt=0.5, t_max=15), dict(dt=2, t_max=40),
                                 dict(dt=4, t_max=60), dict(h=400e3, koz_radius=10.0), dict(rc0=30, wt0=0.0436),
                                 dict(corridor_half_angle=0.2618, h=2000e3)])
def test_sensitivity_axes_step_and_rollout_vs_c_oracle(cfg):
    """The axes of sensitivity_analysis.py:97-134 (dt, h, koz_radius, rc0, wt0, corridor_half_angle): 12 steps of
    512 envs through rdv_step and through rdv_rollout against the C oracle -- other step lengths change how the
    adaptive solver splits the interval (more or fewer RK steps, the clipped last step)."""
    import torch
    from oracle import c_oracle as CO
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv
    from reinforcement_learning_rendezvous_b200.environment_utils import config_to_kwargs
    n, K, seed = 512, 12, 21
    kw = {k: v for k, v in config_to_kwargs(cfg, stochastic=True).items() if v is not None}
    rng = np.random.default_rng(2)
    acts = rng.uniform(-1, 1, (K, n, 6))
    ids, episode = np.arange(n), np.ones(n, dtype=np.int32)
    orc = CO.COracleBatch(CO.make_params(**kw), n)
    orc.reset_from_uniforms(CO.philox_uniforms(seed, ids, episode))
    ref_state, ref_rew = [], []
    for k in range(K):
        _, r, _ = orc.step(acts[k], threads=4)
        ref_state.append(orc.state.copy()); ref_rew.append(r.copy())
    for fused in (False, True):
        env = BatchedRendezvousEnv(n, seed=seed, auto_reset=False, **kw)
        env.reset()
        if fused:
            out = env.rollout(K, actions=torch.as_tensor(acts, device=env.device), record_rewards=True)
            assert rel_err(out["rewards"].cpu().numpy(), np.array(ref_rew)) <= REL_TOL, (cfg, fused)
            assert rel_err(env.get_state().cpu().numpy(), ref_state[-1]) <= REL_TOL, (cfg, fused)
        else:
            for k in range(K):
                _, rew, _ = env.step(torch.as_tensor(acts[k], device=env.device))
                assert rel_err(env.get_state().cpu().numpy(), ref_state[k]) <= REL_TOL, (cfg, k)
                assert rel_err(rew.cpu().numpy(), ref_rew[k]) <= REL_TOL, (cfg, k)
        st = env.read_stats()
        assert st["failures"] == 0 and st["steps"] == n * K


def test_rollout_philox_actions_vs_c_oracle():
    """Device-generated actions + in-warp resets over 60 steps vs the C oracle driven by the same streams."""
    import torch
    from oracle import c_oracle as CO
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv
    n, K, seed, aseed, offset, base = 2048, 60, 11, 0xABCDEF0123, 5000, 1_000_000_000_000
    env = BatchedRendezvousEnv(n, seed=seed, env_offset=offset, t_max=30)
    orc = CO.COracleBatch(CO.make_params(t_max=30), n)
    ids = offset + np.arange(n)
    episode = np.ones(n, dtype=np.int32)
    env.reset()
    orc.reset_from_uniforms(CO.philox_uniforms(seed, ids, episode))
    # two launches of 30 steps: the stream position (step_base) carries over
    outs = [env.rollout(30, action_seed=aseed, step_base=base + 30 * j, record_rewards=True, record_dones=True,
                        record_actions=True, record_obs=True) for j in range(2)]
    rewards = torch.cat([o["rewards"] for o in outs]).cpu().numpy()
    dones = torch.cat([o["dones"] for o in outs]).cpu().numpy()
    actions = torch.cat([o["actions"] for o in outs]).cpu().numpy()
    obs_steps = torch.cat([o["obs_steps"] for o in outs]).cpu().numpy()
    total_done = 0
    for k in range(K):
        a = CO.philox_actions(aseed, ids, base + k)
        np.testing.assert_array_equal(actions[k], a)
        o_obs, o_rew, o_done = orc.step(a, threads=8)
        o_obs, o_rew, o_done = o_obs.copy(), o_rew.copy(), o_done.copy()
        np.testing.assert_array_equal(dones[k], o_done)
        assert rel_err(rewards[k], o_rew) <= REL_TOL
        d = np.flatnonzero(o_done)
        if d.size:
            total_done += d.size
            episode[d] += 1
            mask = np.zeros(n, dtype=np.uint8)
            mask[d] = 1
            o_obs = orc.reset_from_uniforms(CO.philox_uniforms(seed, ids, episode), mask=mask).copy()
        assert np.abs(obs_steps[k] - o_obs).max() <= 1.2e-7
    assert total_done > n
    assert rel_err(env.get_state().cpu().numpy(), orc.state) <= REL_TOL
    np.testing.assert_array_equal(env.episode_index.cpu().numpy(), episode)
    np.testing.assert_array_equal(env.collided.cpu().numpy(), orc.flags[:, 0])
    assert np.abs(env.obs.cpu().numpy() - orc.observe()).max() <= 1.2e-7
    st = env.read_stats()
    assert st["steps"] == n * K and st["episodes"] == total_done and st["failures"] == 0


def test_rollout_shard_invariance_and_edges():
    """Two shards == one batch (global env ids key both streams); n not a multiple of 32; steps = 0; bad args."""
    import torch
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv
    whole = BatchedRendezvousEnv(777, seed=9)
    lo = BatchedRendezvousEnv(400, seed=9, env_offset=0)
    hi = BatchedRendezvousEnv(377, seed=9, env_offset=400)
    for e in (whole, lo, hi):
        e.reset()
        e.rollout(40, action_seed=4)
    s = whole.get_state()
    assert torch.equal(s[:400], lo.get_state()) and torch.equal(s[400:], hi.get_state())
    assert torch.equal(whole.obs[:400], lo.obs) and torch.equal(whole.obs[400:], hi.obs)
    before = whole.get_state().clone()
    whole.rollout(0, action_seed=1)
    assert torch.equal(before, whole.get_state())
    with pytest.raises(ValueError):
        whole.rollout(3)
    with pytest.raises(ValueError):
        whole.rollout(3, actions=torch.zeros((2, 777, 6), dtype=torch.float64, device=whole.device))
    # no auto-reset: finished envs keep stepping, like step()
    a = BatchedRendezvousEnv(64, seed=1, auto_reset=False)
    b = BatchedRendezvousEnv(64, seed=1, auto_reset=False)
    a.reset(); b.reset()
    acts = torch.rand((10, 64, 6), dtype=torch.float64, device=a.device) * 2 - 1
    a.rollout(10, actions=acts)
    for k in range(10):
        b.step(acts[k])
    assert rel_err(a.get_state().cpu().numpy(), b.get_state().cpu().numpy()) <= 1e-11
    assert (a.episode_index == 1).all()


def _tune(key, value):
    from reinforcement_learning_rendezvous_b200 import _native as N
    return N.lib().rdv_tune(key, int(value))


def test_rollout_launch_shapes_are_bit_identical():
    """The lock-step (<= 256 envs per CTA) and the sequential (larger CTAs) solvers, and every CTA size, give the
    same bits: results cannot depend on the batch size or on how a batch is sharded over GPUs."""
    import torch
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv
    n, K = 3000, 40
    states, obs, stats = [], [], []
    from reinforcement_learning_rendezvous_b200 import _native as N
    try:
        for tpb in (256, 384, 448, 512):
            _tune(N.TUNE_ROLLOUT_TPB, tpb)
            env = BatchedRendezvousEnv(n, seed=2, t_max=25)
            env.reset()
            out = env.rollout(K, action_seed=7, record_rewards=True)
            states.append(env.get_state().clone()); obs.append(out["rewards"].clone()); stats.append(env.read_stats())
    finally:
        _tune(N.TUNE_ROLLOUT_TPB, 0)
    for k in range(1, 4):
        assert torch.equal(states[0], states[k]) and torch.equal(obs[0], obs[k])
        assert stats[0]["rk_accepted"] == stats[k]["rk_accepted"] and stats[0]["episodes"] == stats[k]["episodes"]


def test_reset_prefetch_is_bit_identical():
    """The rollout keeps every env's NEXT reset state ready (shared memory inside a launch, a device scratch between
    launches) and refills the used rows every few steps (the reset of (env, episode) does not depend on the
    trajectory).  On-demand resets (period 0), short and long refill periods, rows carried or not carried from
    launch to launch, and short episodes (t_max = 3: envs finish again before the next refill) give the same bits."""
    import torch
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv, _native as N
    try:
        for kw in (dict(t_max=25), dict(t_max=3)):
            ref = None
            for period, carry in ((0, False), (0, True), (2, True), (3, False), (8, True), (8, False), (16, True)):
                _tune(N.TUNE_RESET_REFILL, period)
                env = BatchedRendezvousEnv(2500, seed=5, **kw)
                env.reset()
                out = env.rollout(70, action_seed=3, record_rewards=True, record_dones=True, carry_reset_rows=carry)
                env.rollout(33, action_seed=3, step_base=70, carry_reset_rows=carry)   # a second launch
                got = (env.get_state().clone(), out["rewards"].clone(), out["dones"].clone(), env.episode_index.clone(),
                       env.read_stats())
                if ref is None:
                    ref = got
                    assert int(out["dones"].sum()) > 2500                  # resets really happened
                    continue
                for a, b in zip(ref[:4], got[:4]):
                    assert torch.equal(a, b), (kw, period, carry)
                for key, v in ref[4].items():  # counters exact; the fp64 sums are atomically reduced over CTAs in any order
                    assert got[4][key] == v if float(v).is_integer() else abs(got[4][key] - v) <= 1e-12 * abs(v), key
    finally:
        _tune(N.TUNE_RESET_REFILL, 12)


def test_reset_rows_carried_over_many_short_launches():
    """Rows carried over N short launches == one long launch (bit for bit), also when the caller resets envs, changes
    the seed or overwrites the state between launches, and for the policy-fused variant (rows in the global scratch)."""
    import torch
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv
    n = 3000
    one = BatchedRendezvousEnv(n, seed=8, t_max=12)
    one.reset()
    one.rollout(96, action_seed=2, carry_reset_rows=False)
    for chunk in (1, 5, 16):
        many = BatchedRendezvousEnv(n, seed=8, t_max=12)
        many.reset()
        for k in range(0, 96, chunk):
            many.rollout(min(chunk, 96 - k), action_seed=2, step_base=k)
        assert many.reset_rows is not None and float(many.reset_rows[-1].max()) > 1      # rows are in use
        assert torch.equal(one.get_state(), many.get_state()) and torch.equal(one.i32, many.i32), chunk
        assert torch.equal(one.obs, many.obs)
        sa, sb = one.read_stats(), many.read_stats()
        assert sa["episodes"] == sb["episodes"] and sa["rk_accepted"] == sb["rk_accepted"]
    # interventions between launches: the carried rows must never leak a stale state
    a = BatchedRendezvousEnv(n, seed=8, t_max=12)
    b = BatchedRendezvousEnv(n, seed=8, t_max=12)
    for e, carry in ((a, True), (b, False)):
        e.reset()
        e.rollout(20, action_seed=2, carry_reset_rows=carry)
        mask = torch.zeros(n, dtype=torch.uint8, device=e.device)
        mask[::3] = 1
        e.reset(mask=mask)                                  # bumps the episode index of a third of the envs
        e.rollout(20, action_seed=2, step_base=20, carry_reset_rows=carry)
        e.seed = 99                                         # another reset stream from here on
        e.rollout(20, action_seed=2, step_base=40, carry_reset_rows=carry)
        e.load_state_dict(e.state_dict())
        e.rollout(7, action_seed=2, step_base=60, carry_reset_rows=carry)
    assert torch.equal(a.get_state(), b.get_state()) and torch.equal(a.i32, b.i32)
    # policy-fused variant: rows live in the global scratch
    pol = _policy()
    p1 = BatchedRendezvousEnv(n, seed=4, t_max=10)
    p2 = BatchedRendezvousEnv(n, seed=4, t_max=10)
    p1.reset(); p2.reset()
    p1.rollout(48, policy=pol, carry_reset_rows=False)
    for k in range(0, 48, 12):
        p2.rollout(12, policy=pol)
    assert float(p2.reset_rows[-1].max()) > 1
    assert torch.equal(p1.get_state(), p2.get_state()) and torch.equal(p1.i32, p2.i32)
    assert p1.read_stats()["episodes"] == p2.read_stats()["episodes"] > n


def test_helper_warps_are_bit_identical():
    """65,536 envs on 148 SMs run as 14 worker warps + 2 helper warps per CTA; the helpers recompute used reset rows on
    request and the workers never wait for them (a row that is not ready is computed on the spot).  Same bits as the
    plain kernel: long and short episodes (t_max = 4: envs finish again before their row is back), rows handed from
    launch to launch in both directions (a helper launch leaves not-ready rows tagged as such), caller interventions
    between launches, one SM left free, and forced onto small CTAs."""
    import torch
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv, _native as N

    def run(n, helpers_of_launch, **kw):
        env = BatchedRendezvousEnv(n, seed=3, **kw)
        env.reset()
        base, outs = 0, []
        for j, (steps, carry, reserve) in enumerate(((37, True, 0), (20, True, 0), (5, True, 1), (20, True, 0),
                                                     (9, False, 0), (16, True, 0))):
            _tune(N.TUNE_ROLLOUT_HELPERS, helpers_of_launch(j))
            env.sm_reserve = reserve
            out = env.rollout(steps, action_seed=5, step_base=base, record_rewards=True, record_dones=True,
                              carry_reset_rows=carry)
            outs += [out["rewards"].clone(), out["dones"].clone()]
            base += steps
            if j == 2:
                mask = torch.zeros(n, dtype=torch.uint8, device=env.device)
                mask[::5] = 1
                env.reset(mask=mask)                    # bumps episode indices: carried rows of those envs are stale
        return [env.get_state().clone(), env.i32.clone(), env.obs.clone()] + outs, env.read_stats()

    try:
        for n, kw in ((65536, dict()), (65536, dict(t_max=4)), (62000, dict(t_max=9))):
            ref, ref_stats = run(n, lambda j: 0, **kw)
            assert ref_stats["episodes"] > n
            for pattern in (lambda j: 1, lambda j: j % 2, lambda j: (j + 1) % 2):
                got, got_stats = run(n, pattern, **kw)
                for a, b in zip(ref, got):
                    assert torch.equal(a, b), (n, kw)
                for key in ("steps", "episodes", "rk_accepted", "rk_rejected", "end_attitude", "end_time"):
                    assert ref_stats[key] == got_stats[key], key
        # small batch forced onto the 448-thread shape: mostly idle lanes, helpers on
        _tune(N.TUNE_ROLLOUT_TPB, 448)
        small = []
        for h in (0, 1):
            small.append(run(3000, lambda j: h, t_max=6)[0])
        for a, b in zip(*small):
            assert torch.equal(a, b)
    finally:
        _tune(N.TUNE_ROLLOUT_TPB, 0)
        _tune(N.TUNE_ROLLOUT_HELPERS, 1)


def _policy():
    import os
    from helpers import GOLDEN
    from reinforcement_learning_rendezvous_b200 import MlpPolicy
    return MlpPolicy.load(os.path.join(GOLDEN, "policy.npz"))


def test_fused_policy_rollout_matches_policy_kernel_and_step():
    """Closed-loop rollout with the actor evaluated inside the launch (split-fp16 tensor-core MLP) vs the same loop
    assembled from rdv_policy_forward (fp32 FFMA) + rdv_step.  Actions agree to fp32 rounding; the closed loops
    are compared while that difference has not been amplified (12 steps) and statistically afterwards."""
    import torch
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv
    pol = _policy()
    mc = golden_mc()
    n, K = 1000, 60
    ics = mc["ic_raw"].copy()
    ics[:, 6:10] /= np.linalg.norm(ics[:, 6:10], axis=1, keepdims=True)
    ics[:, 13:17] /= np.linalg.norm(ics[:, 13:17], axis=1, keepdims=True)
    kw = dict(dt=1, t_max=60, rc0_range=0, vc0_range=0, qc0_range=0, wc0_range=0, qt0_range=0, wt0_range=0)
    a = BatchedRendezvousEnv(n, auto_reset=False, **kw)
    b = BatchedRendezvousEnv(n, auto_reset=False, **kw)
    for e in (a, b):
        e.reset()
        e.set_state(ics, reset_counters=False)
    out = a.rollout(K, policy=pol, record_actions=True, record_obs=True, record_rewards=True, record_dones=True)
    obs = b.observe()
    alive = torch.ones(n, dtype=torch.bool, device=b.device)
    max_act = 0.0
    for k in range(K):
        act = pol.forward(obs)
        if k < 12:
            max_act = max(max_act, float((out["actions"][k] - act)[alive].abs().max()))
            assert float((out["obs_steps"][k - 1] - obs)[alive].abs().max()) < 2e-5 if k else True
        obs_k, rew, done = b.step(act)
        obs = obs_k.clone()
        alive &= ~done.bool()
    assert max_act < 2e-5, max_act
    # every recorded action is the actor's output for the observation the kernel saw (open-loop check, all steps)
    prev = torch.cat([a_obs0(ics, kw)[None], out["obs_steps"][:-1]])
    worst = 0.0
    for k in range(0, K, 7):
        worst = max(worst, float((pol.forward(prev[k].contiguous()) - out["actions"][k]).abs().max()))
    assert worst < 5e-6, worst
    # same episodes, statistically: first-done step of every env
    first_a = torch.where(out["dones"].bool().any(0), out["dones"].bool().float().argmax(0), torch.full((n,), K, device=a.device))
    assert (first_a < K).float().mean() > 0.5


def golden_mc():
    from helpers import golden
    return golden("mc.npz")


def a_obs0(ics, kw):
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv
    e = BatchedRendezvousEnv(len(ics), auto_reset=False, **kw)
    e.reset()
    e.set_state(ics, reset_counters=False)
    return e.observe()


def test_fused_policy_monte_carlo_success_rate():
    """The published Monte-Carlo experiment as ONE launch: 1000 initial conditions, 60 closed-loop steps with the
    shipped policy; success and collision counts of the published workbook (545 / 166) within +-5."""
    import torch
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv
    pol = _policy()
    mc = golden_mc()
    ics = mc["ic_raw"].copy()
    ics[:, 6:10] /= np.linalg.norm(ics[:, 6:10], axis=1, keepdims=True)
    ics[:, 13:17] /= np.linalg.norm(ics[:, 13:17], axis=1, keepdims=True)
    env = BatchedRendezvousEnv(1000, auto_reset=False, dt=1, t_max=60, rc0_range=0, vc0_range=0, qc0_range=0,
                               wc0_range=0, qt0_range=0, wt0_range=0)
    env.reset()
    env.set_state(ics, reset_counters=False)
    out = env.rollout(60, policy=pol, record_dones=True)
    done = out["dones"].bool()
    ep_len = torch.where(done.any(0), done.float().argmax(0) + 1, torch.full((1000,), 60, device=env.device))
    same_len = (ep_len.cpu().numpy() == mc["workbook_ep_len"]).mean()
    assert same_len > 0.985, same_len
    st = env.read_stats()
    assert st["failures"] == 0


def test_stochastic_policy_rollout_statistics_and_determinism():
    """Sampling mode: recorded actions = actor mean + exp(log_std) * N(0,1) draws (moments, independence across
    envs / steps / components), the env receives the clipped draw, and the launch is reproducible."""
    import torch
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv
    pol = _policy()
    n, K = 8192, 24
    runs = []
    for _ in range(2):
        env = BatchedRendezvousEnv(n, seed=3)
        obs0 = env.reset().clone()
        out = env.rollout(K, policy=pol, stochastic=True, action_seed=42, step_base=1000, record_actions=True,
                          record_obs=True, record_rewards=True)
        runs.append((obs0, out, env.get_state().clone()))
    (obs0, out, state), (_, out2, state2) = runs
    assert torch.equal(out["actions"], out2["actions"]) and torch.equal(state, state2)
    prev = torch.cat([obs0[None], out["obs_steps"][:-1]])
    std = pol.log_std.exp()
    # actor mean (unclipped) from the fp32 torch weights
    x = prev.reshape(-1, 17)
    h = torch.tanh(x @ pol.w["w0"].T + pol.w["b0"])
    h = torch.tanh(h @ pol.w["w1"].T + pol.w["b1"])
    mean = (h @ pol.w["w2"].T + pol.w["b2"]).reshape(K, n, 6)
    z = ((out["actions"] - mean) / std).double()
    m = z.numel()
    assert abs(float(z.mean())) < 5 / np.sqrt(m) and abs(float(z.var()) - 1) < 0.01
    assert abs(float((z ** 3).mean())) < 0.02 and abs(float((z ** 4).mean()) - 3) < 0.05
    zc = z.reshape(-1, 6)
    corr = torch.corrcoef(zc.T)
    assert float((corr - torch.eye(6, device=corr.device, dtype=corr.dtype)).abs().max()) < 0.01
    assert abs(float(torch.corrcoef(torch.stack([z[0, :, 0], z[1, :, 0]]))[0, 1])) < 0.05
    # a different step_base / seed gives different noise
    env = BatchedRendezvousEnv(n, seed=3)
    env.reset()
    other = env.rollout(K, policy=pol, stochastic=True, action_seed=43, step_base=1000, record_actions=True)
    assert not torch.equal(other["actions"][0], out["actions"][0])
    with pytest.raises(ValueError):
        env.rollout(2, policy=pol, stochastic=True)


def test_sm_reserve_and_ragged_parameter_table_do_not_change_results():
    """Leaving SMs free for a concurrent collective (sm_reserve) only changes how the batch is cut into CTA slices, and a
    parameter table whose last group is not a multiple of 32 envs behaves like separate envs."""
    import torch
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv
    a = BatchedRendezvousEnv(5000, seed=4, t_max=20)
    b = BatchedRendezvousEnv(5000, seed=4, t_max=20)
    b.sm_reserve = 3
    for e in (a, b):
        e.reset()
        e.rollout(45, action_seed=9)
        e.rollout(10, action_seed=9, step_base=45)
    assert torch.equal(a.get_state(), b.get_state()) and torch.equal(a.i32, b.i32) and torch.equal(a.obs, b.obs)
    with pytest.raises(RuntimeError):
        b.sm_reserve = 100000
        b.rollout(1, action_seed=1)
    # ragged table: groups of 64, 32 and 19 envs
    batches = [(64, dict(dt=0.5, t_max=10)), (32, dict(koz_radius=8.0, t_max=10)), (19, dict(h=500e3, t_max=10))]
    env = BatchedRendezvousEnv(115, seed=2, param_batches=batches)
    assert env.param_table is not None
    env.reset()
    out = env.rollout(30, action_seed=5, record_rewards=True, record_dones=True, record_obs=True)
    lo = 0
    for count, kw in batches:
        e = BatchedRendezvousEnv(count, seed=2, env_offset=lo, **kw)
        e.reset()
        o = e.rollout(30, action_seed=5, record_rewards=True, record_dones=True, record_obs=True)
        assert torch.equal(o["rewards"], out["rewards"][:, lo:lo + count])
        assert torch.equal(o["dones"], out["dones"][:, lo:lo + count])
        assert torch.equal(o["obs_steps"], out["obs_steps"][:, lo:lo + count])
        assert torch.equal(e.get_state(), env.get_state()[lo:lo + count])
        lo += count
    # masked reset and the evaluator queries go through the table too
    mask = torch.zeros(115, dtype=torch.uint8, device=env.device)
    mask[60:100] = 1
    env.reset(mask=mask)
    err, col, suc, koz = env.errors()
    assert bool(torch.isfinite(err).all()) and int(env.step_count[60:100].max()) == 0


def test_evaluator_mode_with_tensor_actions_matches_per_step_accumulation():
    """RdvRolloutIO.mc_out without the fused actor (given actions): the device-side accumulators of
    monte_carlo.evaluate against the same quantities accumulated on the host from rdv_step + rdv_errors."""
    import torch
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv, _native as N
    from reinforcement_learning_rendezvous_b200.monte_carlo import _terminal_errors_batch
    n, K = 600, 40
    # starts around the docking point (|rd| = 2 m on the target's -y axis), inside the entry corridor, with a target that
    # barely moves: the four errors cross their limits in both directions, so every level of the first-index rule occurs
    kw = dict(t_max=40, rc0=np.array([0.0, -2.3, 0.0]), rc0_range=0.5, vc0_range=0.05, qt0_range=0.1, wt0_range=0.01)
    rng = np.random.default_rng(3)
    acts = torch.as_tensor(rng.uniform(-0.3, 0.3, (K, n, 6)), device="cuda")
    a = BatchedRendezvousEnv(n, seed=1, auto_reset=False, **kw)
    b = BatchedRendezvousEnv(n, seed=1, auto_reset=False, **kw)
    a.reset(); b.reset()
    mc = a.rollout(K, actions=acts, monte_carlo=True)["mc"].cpu().numpy()
    err, col, suc, koz = b.errors()
    errs = np.full((K + 1, n, 4), np.nan)
    errs[0] = err.cpu().numpy()
    n_col, n_suc = col.cpu().numpy().astype(int), suc.cpu().numpy().astype(int)
    min_koz, total, length = koz.cpu().numpy().copy(), np.zeros(n), np.zeros(n, dtype=int)
    alive = np.ones(n, dtype=bool)
    tdv = np.zeros(n)
    for k in range(K):
        _, rew, done = b.step(acts[k])
        err, col, suc, koz = (t.cpu().numpy() for t in b.errors())
        errs[k + 1][alive] = err[alive]
        n_col += (col.astype(bool) & alive)
        n_suc += (suc.astype(bool) & alive)
        min_koz = np.where(alive & (koz < min_koz), koz, min_koz)
        total += np.where(alive, rew.cpu().numpy(), 0.0)
        length += alive
        fin = alive & done.cpu().numpy().astype(bool)
        tdv = np.where(fin, b.total_delta_v.cpu().numpy(), tdv)
        alive &= ~done.cpu().numpy().astype(bool)
    p = b.params
    te = _terminal_errors_batch(errs, length, (p.max_rd_error, p.max_vd_error, p.max_qd_error, p.max_wd_error))
    np.testing.assert_array_equal(mc[:, N.MC_EP_LEN], length)
    np.testing.assert_array_equal(mc[:, N.MC_NUM_COLLISIONS], n_col)
    np.testing.assert_array_equal(mc[:, N.MC_NUM_SUCCESSES], n_suc)
    assert rel_err(mc[:, N.MC_MIN_KOZ], min_koz) <= 1e-10 and rel_err(mc[:, N.MC_TOTAL_REWARD], total) <= 1e-10
    done_rows = mc[:, N.MC_END_REASON] >= 0
    assert done_rows.sum() > n // 2
    assert rel_err(mc[done_rows, N.MC_TOTAL_DELTA_V], tdv[done_rows]) <= 1e-12
    got = mc[:, [N.MC_POS_ERR, N.MC_VEL_ERR, N.MC_ATT_ERR, N.MC_ROT_ERR]].copy()
    got[:, 2:] = np.degrees(got[:, 2:])
    assert np.max(np.abs(got - te) / np.maximum(np.abs(te), 1e-3)) <= 1e-9
    assert len(np.unique(mc[:, N.MC_LEVEL])) >= 3                  # several branches of the first-index rule ran
    assert mc[:, N.MC_NUM_SUCCESSES].max() > 0 and mc[:, N.MC_TAIL_COUNT].max() > 1
