"""Host-side scaling of RendezvousVecEnv.step over the GPUs of one box (development tool).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        tools/e2e_multi.py [n_envs] [steps]

Every rank drives its own GPU through the VecEnv drop-in, all ranks at the same time (gloo barrier between the modes):
per-phase host time of a step without any CPU binding, with the rank bound to the cores NVML reports as local to its
GPU (`distributed.bind_to_gpu_cpus`), and printed next to the box's CPU / NUMA topology.
"""
import gc
import glob
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reinforcement_learning_rendezvous_b200 import RendezvousVecEnv                     # noqa: E402
from reinforcement_learning_rendezvous_b200.distributed import bind_to_gpu_cpus, gpu_cpu_affinity  # noqa: E402


def topology(local_rank):
    out = {"cpu_count": os.cpu_count(), "sched_affinity": len(os.sched_getaffinity(0))}
    for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        if os.path.exists(f):
            out[f] = open(f).read().strip()
    out["numa"] = {os.path.basename(os.path.dirname(p)): open(p).read().strip()
                   for p in sorted(glob.glob("/sys/devices/system/node/node*/cpulist"))}
    out["gpu_affinity"] = sorted(gpu_cpu_affinity(local_rank))
    return out


def run(n, steps, local_rank):
    v = RendezvousVecEnv(n, device=f"cuda:{local_rank}", seed=local_rank)
    v.reset()
    rng = np.random.default_rng(local_rank)
    ring = rng.uniform(-1, 1, (8, n, 6)).astype(np.float32)
    for k in range(30):
        v.step(ring[k % 8])
    dist.barrier()
    T = dict(stage=0.0, launch_fetch=0.0, infos=0.0, total=0.0)
    for k in range(steps):
        t0 = time.perf_counter()
        v.step_async(ring[k % 8])
        t1 = time.perf_counter()
        gc.disable()
        obs, rew, done, rows = v._launch_and_fetch(renew_infos=True)
        t2 = time.perf_counter()
        v._build_infos(rows)
        gc.enable()
        del rows
        t3 = time.perf_counter()
        T["stage"] += t1 - t0; T["launch_fetch"] += t2 - t1; T["infos"] += t3 - t2; T["total"] += t3 - t0
    dist.barrier()
    t0 = time.perf_counter()
    for k in range(steps):
        v.step(ring[k % 8])
    whole = (time.perf_counter() - t0) / steps
    dist.barrier()
    t0 = time.perf_counter()
    for k in range(steps):
        v.step_arrays(ring[k % 8])
    arrays = (time.perf_counter() - t0) / steps
    res = {k: round(1e3 * x / steps, 4) for k, x in T.items()}
    res["step"] = round(1e3 * whole, 4)
    res["step_arrays"] = round(1e3 * arrays, 4)
    del v
    gc.collect()
    return res


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist.init_process_group("gloo")
    torch.cuda.set_device(local_rank)
    topo = topology(local_rank)
    tops = [None] * world
    dist.all_gather_object(tops, topo)
    if local_rank == 0:
        print("topology:", {k: v for k, v in topo.items() if k != "gpu_affinity"}, flush=True)
        for r, t in enumerate(tops):
            a = t["gpu_affinity"]
            print(f"  gpu {r}: {len(a)} local cpus {a[:4]}..{a[-2:] if a else a}", flush=True)
    full = os.sched_getaffinity(0)
    for mode in ("unbound", "bound", "unbound", "bound"):
        os.sched_setaffinity(0, full)
        cores = None
        if mode == "bound":
            cores = bind_to_gpu_cpus(local_rank, local_rank, world)
        r = run(n, steps, local_rank)
        allr = [None] * world
        dist.all_gather_object(allr, (r, sorted(cores) if cores else None))
        if local_rank == 0:
            tot = sum(n / (x[0]["step"] * 1e-3) for x in allr)
            tot_a = sum(n / (x[0]["step_arrays"] * 1e-3) for x in allr)
            print(f"{mode}: {world} ranks, whole box {tot / 1e6:.1f} M env-steps/s through step(), "
                  f"{tot_a / 1e6:.1f} M through step_arrays()", flush=True)
            for k, x in enumerate(allr):
                print(f"  rank {k}: {x[0]} cores {x[1]}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
