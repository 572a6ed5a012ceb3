"""A/B timing of the step / reset kernels for alternative builds of librdv_b200.so (development tool).

    python tools/ab_step.py build/lib_a.so build/lib_b.so ...

Each library is timed in its own subprocess (RDV_B200_LIB) on the same GPU, interleaved twice, with CUDA events
around every launch and the back-to-back step rate.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys, json, torch
sys.path.insert(0, %r)
from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv
n, R = int(sys.argv[1]), 300
env = BatchedRendezvousEnv(n, seed=0, integrator=sys.argv[2])
env.reset()
g = torch.Generator(device='cuda'); g.manual_seed(1)
ring = torch.rand((16, n, 6), dtype=torch.float64, device='cuda', generator=g) * 2 - 1
for k in range(30): env.step(ring[k %% 16])
torch.cuda.synchronize()
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(R)]
for k in range(R):
    ev[k][0].record(); env.step(ring[k %% 16]); ev[k][1].record()
torch.cuda.synchronize()
ts = sorted(a.elapsed_time(b) for a, b in ev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(1000): env.step(ring[k %% 16])
e1.record(); torch.cuda.synchronize()
print(json.dumps(dict(step_med_us=1e3 * ts[R // 2], step_min_us=1e3 * ts[0], 
                      loop_us=e0.elapsed_time(e1), rk=env.read_stats()['rk_accepted'] / max(env.read_stats()['steps'], 1) / 2)))
""" % ROOT


def main():
    libs = [a for a in sys.argv[1:] if not a.startswith("--")]
    n = 65536
    integ = "rk45"
    for a in sys.argv[1:]:
        if a.startswith("--n="):
            n = int(a[4:])
        if a.startswith("--integrator="):
            integ = a.split("=", 1)[1]
    for rep in range(2):
        for lib in libs:
            env = dict(os.environ, RDV_B200_LIB=os.path.abspath(lib))
            out = subprocess.run([sys.executable, "-c", CHILD, str(n), integ], env=env, capture_output=True, text=True)
            line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:]
            print(f"{os.path.basename(lib):28s} {line}", flush=True)


if __name__ == "__main__":
    main()
