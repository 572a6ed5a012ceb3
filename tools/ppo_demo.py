"""Config #3 demo: PPO (main.py's network and hyper-parameters) with 16,384 GPU envs (development tool).

    python tools/ppo_demo.py [n] [iters] [vecenv|fused|stepwise] [batch_size] [n_epochs]

`vecenv` (default) collects through the SB3 VecEnv drop-in (RendezvousVecEnv.step: numpy actions in, numpy
observations / rewards / dones out) exactly like SB3's collect_rollouts; `fused` collects with one policy-fused
rollout launch per iteration; `stepwise` with one rdv_step launch per step on device tensors.  Collection and update
rates are reported separately.
"""
import json
import sys
sys.path.insert(0, '.')
from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv, RendezvousVecEnv
from reinforcement_learning_rendezvous_b200.ppo import PPO, PPOConfig

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
mode = sys.argv[3] if len(sys.argv) > 3 else "vecenv"
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 16384
epochs = int(sys.argv[5]) if len(sys.argv) > 5 else 10
env = RendezvousVecEnv(n, seed=0) if mode == "vecenv" else BatchedRendezvousEnv(n, seed=0)
cfg = PPOConfig(n_steps=16, batch_size=batch, n_epochs=epochs, n_evals=256, fused=(mode == "fused"))
algo = PPO(env, cfg)
algo.learn(iters * cfg.n_steps * n, eval_every=5, verbose=False)
print(json.dumps(dict(mode=mode, envs=n, n_steps=cfg.n_steps, batch_size=batch, n_epochs=epochs, iterations=iters)))
for r in cfg.log[::max(1, iters // 10)] + cfg.log[-1:]:
    print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()}))
rows = cfg.log[2:]
print(json.dumps(dict(
    collect_env_steps_per_s=round(sum(cfg.n_steps * n for _ in rows) / sum(r["collect_s"] for r in rows), 1),
    update_samples_per_s=round(sum(cfg.n_steps * n * epochs for _ in rows) / sum(r["update_s"] for r in rows), 1),
    collect_s_per_iter=round(sum(r["collect_s"] for r in rows) / len(rows), 4),
    update_s_per_iter=round(sum(r["update_s"] for r in rows) / len(rows), 4))))
print("final eval", algo.evaluate())
