"""A/B timing of the fused rollout kernel for alternative builds of librdv_b200.so (development tool).

    python tools/ab_rollout.py build/lib_a.so build/lib_b.so ... [--n=65536] [--policy]

Each library is timed in its own subprocess (RDV_B200_LIB) on the same GPU, interleaved twice: short (20-step) and
long (250-step) launches with Philox actions, CUDA events around the batch of launches.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys, json, torch
sys.path.insert(0, %r)
from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv, MlpPolicy
n, policy = int(sys.argv[1]), sys.argv[2] == "1"
env = BatchedRendezvousEnv(n, seed=0)
env.reset()
kw = dict(policy=MlpPolicy.load(%r + "/tests/golden/policy.npz")) if policy else dict(action_seed=1)
env.rollout(64, **kw)
res = {}
base = 64
for K, reps in ((20, 60), (250, 6)):
    env.rollout(K, step_base=base, **kw); base += K
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        env.rollout(K, step_base=base, **kw); base += K
    e1.record(); torch.cuda.synchronize()
    res["us_per_step_K%%d" %% K] = round(1e3 * e0.elapsed_time(e1) / (reps * K), 3)
st = env.read_stats()
res["rk"] = round(st["rk_accepted"] / max(st["steps"], 1) / 2, 6)
res["mean_len"] = round(st["length_sum"] / max(st["episodes"], 1), 4)
print(json.dumps(res))
""" % (ROOT, ROOT)


def main():
    libs = [a for a in sys.argv[1:] if not a.startswith("--")]
    n, policy = 65536, "0"
    for a in sys.argv[1:]:
        if a.startswith("--n="):
            n = int(a[4:])
        if a == "--policy":
            policy = "1"
    for rep in range(2):
        for spec in libs:
            lib, *sets = spec.split(":")                      # build/x.so:RDV_RESET_REFILL=16
            env = dict(os.environ, RDV_B200_LIB=os.path.abspath(lib))
            env.update(dict(kv.split("=", 1) for kv in sets))
            out = subprocess.run([sys.executable, "-c", CHILD, str(n), policy], env=env, capture_output=True, text=True)
            line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-600:]
            print(f"{os.path.basename(spec):44s} {line}", flush=True)


if __name__ == "__main__":
    main()
