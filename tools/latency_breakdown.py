"""Single-CTA latency of one env step under runtime options (development tool): graph-replayed, n = 64 envs."""
import sys, torch
sys.path.insert(0, '.')
from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv

def run(n, **kw):
    env = BatchedRendezvousEnv(n, seed=0, **kw)
    env.reset()
    ring = torch.rand((16, n, 6), dtype=torch.float64, device='cuda') * 2 - 1
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for k in range(16): env.step(ring[k])
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for k in range(16): env.step(ring[k])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.replay(); torch.cuda.synchronize()
    e0.record()
    for _ in range(100): g.replay()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / 1600

for n in (64, 148 * 4 * 64, 65536):
    print(f"n={n}")
    print("  full                         %.2f us" % run(n))
    print("  no auto-reset                %.2f us" % run(n, auto_reset=False))
    print("  no stats                     %.2f us" % run(n, track_stats=False))
    print("  no reset, no stats           %.2f us" % run(n, auto_reset=False, track_stats=False))
    print("  closed form                  %.2f us" % run(n, integrator="closed_form"))
    print("  closed form, no reset/stats  %.2f us" % run(n, integrator="closed_form", auto_reset=False, track_stats=False))
