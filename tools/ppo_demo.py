"""Config #3 demo: PPO (main.py hyper-parameters) with 16,384 GPU envs, device-tensor loop (development tool)."""
import sys, json
sys.path.insert(0, '.')
from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv
from reinforcement_learning_rendezvous_b200.ppo import PPO, PPOConfig
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
env = BatchedRendezvousEnv(n, seed=0)
cfg = PPOConfig(n_steps=16, batch_size=16384, n_epochs=10, n_evals=256)
algo = PPO(env, cfg)
algo.learn(iters * cfg.n_steps * n, eval_every=5, verbose=False)
for r in cfg.log[::max(1, iters // 10)] + cfg.log[-1:]:
    print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()}))
print("final eval", algo.evaluate())
