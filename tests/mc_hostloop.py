"""TEST CHECKER: the Monte-Carlo evaluation assembled from separate launches (rdv_policy_forward + rdv_step +
rdv_errors per step) with the per-episode reduction of monte_carlo.py:159-189 done in numpy on the recorded per-step
arrays.  This was the product path of round 1; the product now does all of it inside one rollout launch
(evaluate_batch), and this loop is what that launch is compared with."""
import numpy as np
import torch


def evaluate_batch_hostloop(policy, initial_states, config=None, reward_kwargs=None, device="cuda"):
    from reinforcement_learning_rendezvous_b200.batched_env import BatchedRendezvousEnv
    from reinforcement_learning_rendezvous_b200.environment_utils import config_to_kwargs
    from reinforcement_learning_rendezvous_b200.monte_carlo import _terminal_errors_batch
    ics = np.array(initial_states, dtype=np.float64, copy=True).reshape(-1, 20)
    ics[:, 6:10] /= np.linalg.norm(ics[:, 6:10], axis=1, keepdims=True)
    ics[:, 13:17] /= np.linalg.norm(ics[:, 13:17], axis=1, keepdims=True)
    m = ics.shape[0]
    kw = config_to_kwargs(dict(dt=1, t_max=60) if config is None else config, stochastic=False)
    env = BatchedRendezvousEnv(m, device=device, auto_reset=False, track_stats=False, reward_kwargs=reward_kwargs, **kw)
    p = env.params
    steps_max = int(p.t_max / p.dt) + 1
    env.reset()
    env.set_state(ics, reset_counters=False)
    obs = env.observe()
    dev = env.device
    err_log = torch.full((steps_max + 1, m, 4), float("nan"), dtype=torch.float64, device=dev)
    alive = torch.ones(m, dtype=torch.bool, device=dev)
    length = torch.zeros(m, dtype=torch.int64, device=dev)
    n_col = torch.zeros(m, dtype=torch.int64, device=dev)
    n_suc = torch.zeros(m, dtype=torch.int64, device=dev)
    total_reward = torch.zeros(m, dtype=torch.float64, device=dev)
    tdv = torch.zeros(m, dtype=torch.float64, device=dev)
    err, col, suc, koz = env.errors()
    err_log[0] = err
    n_col += col.long()
    n_suc += suc.long()
    min_koz = koz.clone()
    actions = torch.empty((m, 6), dtype=torch.float32, device=dev)
    k = 0
    while bool(alive.any()) and k < steps_max:
        k += 1
        policy.forward(obs, out=actions)
        obs, rew, done = env.step(actions)
        err, col, suc, koz = env.errors()
        err_log[k][alive] = err[alive]
        n_col += (col.bool() & alive).long()
        n_suc += (suc.bool() & alive).long()
        min_koz = torch.where(alive & (koz < min_koz), koz, min_koz)
        total_reward += torch.where(alive, rew, torch.zeros_like(rew))
        length += alive.long()
        finished = alive & done.bool()
        tdv = torch.where(finished, env.total_delta_v, tdv)
        alive = alive & ~done.bool()
    length_np = length.cpu().numpy()
    limits = (p.max_rd_error, p.max_vd_error, p.max_qd_error, p.max_wd_error)
    te = _terminal_errors_batch(err_log.cpu().numpy(), length_np, limits)
    n_col_np, n_suc_np = n_col.cpu().numpy(), n_suc.cpu().numpy()
    return dict(
        ep_len=np.round(length_np * p.dt, 3), num_collisions=n_col_np, collided=(n_col_np > 0).astype(np.int64),
        total_reward=total_reward.cpu().numpy(), total_delta_v=tdv.cpu().numpy(), num_successes=n_suc_np,
        succeeded=(n_suc_np > 0).astype(np.int64), min_dist_from_koz=min_koz.cpu().numpy(),
        pos_error=te[:, 0], vel_error=te[:, 1], att_error=te[:, 2], rot_error=te[:, 3])
