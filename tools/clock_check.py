"""SM clock / power while the rollout kernel runs for a few seconds (development tool)."""
import subprocess, sys, time, torch
sys.path.insert(0, '.')
from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
env = BatchedRendezvousEnv(n, seed=0)
env.reset()
env.rollout(64, action_seed=1); torch.cuda.synchronize()
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active,temperature.gpu",
                      "--format=csv,noheader", "-lms", "200"], stdout=subprocess.PIPE, text=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
steps = 0
t0 = time.time()
while time.time() - t0 < 4.0:
    env.rollout(256, action_seed=1, step_base=steps); steps += 256
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
p.terminate()
out = p.stdout.read().strip().splitlines()
print("us/step", 1e3 * e0.elapsed_time(e1) / steps)
print("\n".join(out[::3]))
