"""tcgen05 policy kernel vs the fp32-FMA kernel and torch fp64 (development tool / first-run check)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from reinforcement_learning_rendezvous_b200 import MlpPolicy
pol = MlpPolicy.load("tests/golden/policy.npz")
g = torch.Generator(device='cuda'); g.manual_seed(0)
for n in (1, 100, 128, 129, 1000, 65536, 131072):
    obs = torch.rand((n, 17), dtype=torch.float32, device='cuda', generator=g) * 2 - 1
    a_tc = pol.forward(obs)
    a_ff = pol.forward(obs, ffma=True)
    torch.cuda.synchronize()
    w = {k: v.double() for k, v in pol.w.items()}
    h = torch.tanh(obs.double() @ w["w0"].T + w["b0"]); h = torch.tanh(h @ w["w1"].T + w["b1"])
    ref = torch.clamp(h @ w["w2"].T + w["b2"], -1, 1)
    print(f"n={n:7d}  tc-ffma {float((a_tc - a_ff).abs().max()):.3e}  tc-fp64 {float((a_tc.double() - ref).abs().max()):.3e}  ffma-fp64 {float((a_ff.double() - ref).abs().max()):.3e}", flush=True)
obs = torch.rand((131072, 17), dtype=torch.float32, device='cuda', generator=g) * 2 - 1
for name, kw in (("tcgen05", {}), ("ffma", dict(ffma=True))):
    pol.forward(obs, **kw); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): pol.forward(obs, **kw)
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {1e3 * e0.elapsed_time(e1) / 50:.1f} us for 131072 envs")
