"""Host-side cost of one env.step() call vs GPU time per step (development tool)."""
import sys, time, torch
sys.path.insert(0, '.')
from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv
for n in (64, 65536):
    env = BatchedRendezvousEnv(n, seed=0)
    env.reset()
    ring = torch.rand((16, n, 6), dtype=torch.float64, device='cuda') * 2 - 1
    for k in range(50): env.step(ring[k % 16])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(2000): env.step(ring[k % 16])
    t_submit = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_total = time.perf_counter() - t0
    # CUDA graph of 16 steps
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
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
    print(f"n={n}: submit {1e6*t_submit/2000:.1f} us/step, total {1e6*t_total/2000:.1f} us/step, graph {1e3*e0.elapsed_time(e1)/1600:.2f} us/step")
