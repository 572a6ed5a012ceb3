"""Helper-warp variant of the rollout kernel (rdv_tune RDV_TUNE_ROLLOUT_HELPERS) against the plain one: same bits, and
its timing (development tool)."""
import sys, torch
sys.path.insert(0, '.')
from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv, _native as N
lib = N.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
res = {}
for mode in (0, 1):
    lib.rdv_tune(N.TUNE_ROLLOUT_HELPERS, mode)
    for kw in (dict(), dict(t_max=4)):
        env = BatchedRendezvousEnv(n, seed=3, **kw)
        env.reset()
        out = env.rollout(37, action_seed=5, record_rewards=True, record_dones=True)
        for j in range(3):
            env.rollout(20, action_seed=5, step_base=37 + 20 * j)
        env.rollout(9, action_seed=5, step_base=97, carry_reset_rows=False)
        torch.cuda.synchronize()
        res[(mode, tuple(kw))] = (env.get_state().clone(), env.i32.clone(), out["rewards"].clone(), out["dones"].clone(),
                                  env.obs.clone(), env.read_stats())
for kw in (tuple(), ("t_max",)):
    a, b = res[(0, kw)], res[(1, kw)]
    ok = all(torch.equal(x, y) for x, y in zip(a[:5], b[:5]))
    same_counts = all(a[5][k] == b[5][k] for k in ("steps", "episodes", "rk_accepted", "end_attitude", "end_time"))
    print("helpers vs plain", kw or "default", "identical:", ok, same_counts, "episodes", a[5]["episodes"])
    assert ok and same_counts
for mode in (0, 1):
    lib.rdv_tune(N.TUNE_ROLLOUT_HELPERS, mode)
    env = BatchedRendezvousEnv(n, seed=0)
    env.reset()
    env.rollout(64, action_seed=1)
    for K, reps in ((20, 60), (250, 6)):
        base = 1000
        env.rollout(K, action_seed=1, step_base=base)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(reps):
            base += K
            env.rollout(K, action_seed=1, step_base=base)
        e1.record(); torch.cuda.synchronize()
        print(f"helpers={mode} K={K}: {1e3 * e0.elapsed_time(e1) / (reps * K):.3f} us per step")
lib.rdv_tune(N.TUNE_ROLLOUT_HELPERS, 0)
