"""GPU: the device-tensor PPO loop (main.py's training, without SB3) runs, learns, evaluates through the fused
policy rollout and exports a model that MlpPolicy loads."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fused", [True, False])
def test_ppo_short_training_run(tmp_path, fused):
    import torch
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv, MlpPolicy
    from reinforcement_learning_rendezvous_b200.ppo import PPO, PPOConfig
    env = BatchedRendezvousEnv(4096, seed=0)
    cfg = PPOConfig(n_steps=16, batch_size=8192, n_epochs=4, n_evals=64, seed=0, fused=fused)
    algo = PPO(env, cfg)
    algo.learn(total_timesteps=12 * 16 * 4096, eval_every=4)
    log = cfg.log
    assert len(log) == 12 and all(np.isfinite(r["value_loss"]) and np.isfinite(r["pg_loss"]) for r in log)
    assert algo.num_timesteps == 12 * 16 * 4096
    # the attitude-keeping term dominates early learning: the mean step reward must rise
    first, last = np.mean([r["mean_step_reward"] for r in log[:2]]), np.mean([r["mean_step_reward"] for r in log[-2:]])
    assert last > first + 0.02, (first, last)
    assert max(r["collect_steps_per_s"] for r in log) > 1e6          # the first iteration pays the CUDA warm-up
    ev = algo.evaluate()
    assert 0 < ev["mean_length"] <= 120 and np.isfinite(ev["mean_return"])
    assert algo.best_state is not None
    path = str(tmp_path / "model.npz")
    algo.save(path)
    pol = MlpPolicy.load(path)
    obs = env.obs[:16].clone()
    mean, _ = algo.policy(obs)
    algo.policy.load_state_dict(algo.best_state)
    mean_best, _ = algo.policy(obs)
    act = pol.forward(obs)
    assert (act - mean_best.clamp(-1, 1)).abs().max().item() < 1e-5


def test_ppo_through_the_vecenv_drop_in():
    """BASELINE.json configs[2]: collection through RendezvousVecEnv.step (numpy in / out, the SB3 surface), the same
    update; the buffer rows are what the VecEnv returned and learning still happens."""
    import torch
    from reinforcement_learning_rendezvous_b200 import RendezvousVecEnv
    from reinforcement_learning_rendezvous_b200.ppo import PPO, PPOConfig
    venv = RendezvousVecEnv(2048, seed=0)
    cfg = PPOConfig(n_steps=16, batch_size=4096, n_epochs=4, n_evals=64, seed=0)
    algo = PPO(venv, cfg)
    algo.learn(total_timesteps=10 * 16 * 2048, eval_every=0)
    log = cfg.log
    assert len(log) == 10 and all(np.isfinite(r["value_loss"]) for r in log)
    first, last = np.mean([r["mean_step_reward"] for r in log[:2]]), np.mean([r["mean_step_reward"] for r in log[-2:]])
    assert last > first + 0.01, (first, last)
    st = venv.read_stats()
    assert st["steps"] == 10 * 16 * 2048 and st["episodes"] > 0


def test_captured_update_equals_eager_update():
    """PPO.update with the minibatch step replayed as a CUDA graph == the same update run eagerly: same buffer, same
    permutations -> the same parameters (the capture's warm-up steps leave no trace), also with a ragged last
    minibatch and across two consecutive updates (Adam state carried by the graph)."""
    import torch
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv
    from reinforcement_learning_rendezvous_b200.ppo import PPO, PPOConfig
    algos = []
    for graph in (True, False):
        env = BatchedRendezvousEnv(1000, seed=3)
        cfg = PPOConfig(n_steps=8, batch_size=3000, n_epochs=3, n_evals=8, seed=1, fused=True, cuda_graph=graph)
        algos.append(PPO(env, cfg))
    a, b = algos
    b.policy.load_state_dict(a.policy.state_dict())
    init = [v.clone() for v in a.policy.state_dict().values()]
    for it in range(2):
        adv, ret = a.collect()
        for name in ("buf_obs", "buf_act", "buf_logp", "buf_val", "buf_rew", "buf_done"):
            getattr(b, name).copy_(getattr(a, name))
        torch.manual_seed(100 + it)
        sa = a.update(adv, ret)
        torch.manual_seed(100 + it)
        sb = b.update(adv.clone(), ret.clone())
        assert a._graph is not None and b._graph is None
        for (k, pa), (_, pb) in zip(a.policy.state_dict().items(), b.policy.state_dict().items()):
            assert torch.allclose(pa, pb, rtol=1e-5, atol=1e-6), (it, k, (pa - pb).abs().max().item())
        assert abs(sa["value_loss"] - sb["value_loss"]) <= 1e-4 * max(1.0, abs(sb["value_loss"]))
    assert any((pa - pi).abs().max().item() > 1e-4 for pa, pi in zip(a.policy.state_dict().values(), init))
