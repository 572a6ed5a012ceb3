"""Large parity run: the benchmark workload itself (65,536 envs, Philox actions, in-kernel resets, 250-step launches
with the reset prefetch) checked step by step against the C oracle driven by the same action / reset streams
(a checker script, not collected by pytest; the tests do the same at sizes that finish in seconds).

usage (from the repo root): python tests/large_parity.py [envs] [launches] [steps_per_launch]
"""
import sys, time, json
import numpy as np
import torch
sys.path.insert(0, '.')
from oracle import c_oracle as CO
from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 4
K = int(sys.argv[3]) if len(sys.argv) > 3 else 250
seed, aseed, offset, threads = 5, 0x1234567, 10_000, 16
env = BatchedRendezvousEnv(n, seed=seed, env_offset=offset)
orc = CO.COracleBatch(CO.make_params(), n)
ids = offset + np.arange(n)
episode = np.ones(n, dtype=np.int32)
env.reset()
orc.reset_from_uniforms(CO.philox_uniforms(seed, ids, episode))
worst = dict(reward=0.0, state=0.0, obs=0.0)
total_done = flag_mismatch = 0
t0 = time.time()
for j in range(launches):
    out = env.rollout(K, action_seed=aseed, step_base=j * K, record_rewards=True, record_dones=True, record_obs=True)
    rewards, dones, obs_steps = (out[k].cpu().numpy() for k in ("rewards", "dones", "obs_steps"))
    for k in range(K):
        a = CO.philox_actions(aseed, ids, j * K + k)
        o_obs, o_rew, o_done = orc.step(a, threads=threads)
        o_obs, o_rew, o_done = o_obs.copy(), o_rew.copy(), o_done.copy()
        flag_mismatch += int((dones[k] != o_done).sum())
        worst["reward"] = max(worst["reward"], float(np.max(np.abs(rewards[k] - o_rew) / np.maximum(np.abs(o_rew), 1.0))))
        d = np.flatnonzero(o_done)
        if d.size:
            total_done += d.size
            episode[d] += 1
            mask = np.zeros(n, dtype=np.uint8)
            mask[d] = 1
            o_obs = orc.reset_from_uniforms(CO.philox_uniforms(seed, ids, episode), mask=mask).copy()
        worst["obs"] = max(worst["obs"], float(np.abs(obs_steps[k] - o_obs).max()))
    st = env.get_state().cpu().numpy()
    worst["state"] = max(worst["state"], float(np.max(np.abs(st - orc.state) / np.maximum(np.abs(orc.state), 1.0))))
    print(f"launch {j}: {n * K * (j + 1):,} env-steps checked, {total_done:,} episodes, done-flag mismatches "
          f"{flag_mismatch}, worst rel dev reward {worst['reward']:.2e} state {worst['state']:.2e} obs {worst['obs']:.2e} "
          f"({time.time() - t0:.0f} s)", flush=True)
stats = env.read_stats()
ok = flag_mismatch == 0 and worst["reward"] <= 1e-9 and worst["state"] <= 1e-9 and worst["obs"] <= 1.2e-7 \
    and np.array_equal(env.episode_index.cpu().numpy(), episode) and stats["failures"] == 0
print(json.dumps(dict(envs=n, steps=launches * K, env_steps=n * launches * K, episodes=int(total_done),
                      done_flag_mismatches=flag_mismatch, worst_rel_dev=worst, rk_steps_per_solve=stats["rk_accepted"] / (2 * stats["steps"]),
                      kernel_failures=stats["failures"], ok=bool(ok))))
sys.exit(0 if ok else 1)
