"""Multi-GPU sharding of the environment batch.

Every environment is independent, so the step path needs NO collective: rank g owns the
contiguous global env range ``shard_range(total, world, g)`` and resets are keyed by the global
env id, which makes results invariant to the number of GPUs.  The only exchange is the
per-rollout statistics vector (16 doubles), summed with one all-reduce -- NCCL over NVLink on the
GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

import os
from typing import Optional, Set, Tuple

import torch
import torch.distributed as dist

from . import _native as N


def shard_range(total_envs: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced partition: the first ``total % world`` ranks get one extra env."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(int(total_envs), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def make_sharded_env(total_envs: int, rank: Optional[int] = None, world_size: Optional[int] = None, device=None,
                     seed: int = 0, **kwargs):
    """This rank's shard of a ``total_envs`` batch as a BatchedRendezvousEnv (env_offset = shard start)."""
    from .batched_env import BatchedRendezvousEnv
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(total_envs, world_size, rank)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return BatchedRendezvousEnv(hi - lo, device=device, seed=seed, env_offset=lo, **kwargs)


def gpu_cpu_affinity(device_index: int) -> Set[int]:
    """The host cores NVML reports as local to a GPU (same socket / PCIe root), intersected with the cores this
    process may run on.  Empty when NVML cannot say (no driver, no permission)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = int(device_index)
        if visible:
            ids = [x.strip() for x in visible.split(",") if x.strip()]
            if index < len(ids) and ids[index].isdigit():
                index = int(ids[index])
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cores = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
    except Exception:
        return set()
    return cores & set(os.sched_getaffinity(0))


def bind_to_gpu_cpus(device_index: int, local_rank: int = 0, local_world: int = 1) -> Optional[Set[int]]:
    """One process per GPU: keep this process (the Python thread that stages actions and builds the per-episode
    objects of a VecEnv step, and the pinned buffers it first touches) on host cores local to its GPU.  The GPU's
    local cores are dealt out among the ranks whose GPUs share them, so that ranks do not sit on each other's cores or
    on the sibling hyper-threads of a busy neighbour more than the box forces them to.  Returns the cores it bound to,
    or None when NVML gave no answer (nothing is changed then)."""
    mine = gpu_cpu_affinity(device_index)
    if not mine:
        return None
    # ranks with the same local core set share it: rank r takes every k-th core starting at its position among them
    sharers = [r for r in range(int(local_world)) if gpu_cpu_affinity(r) == mine] or [int(local_rank)]
    pos = sharers.index(int(local_rank)) if int(local_rank) in sharers else 0
    ordered = sorted(mine)
    per = max(1, len(ordered) // len(sharers))
    cores = set(ordered[pos * per:(pos + 1) * per]) or mine
    os.sched_setaffinity(0, cores)
    return cores


def all_reduce_stats(stats: torch.Tensor, group=None, async_op: bool = False):
    """Sum the [RDV_NSTATS] statistics vector over all ranks in place.  No-op without a process group."""
    if stats.numel() != N.NSTATS:
        raise ValueError(f"stats must have {N.NSTATS} elements")
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    return dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


class OverlappedStatsReducer:
    """Per-rollout statistics summed over ranks with the collective of rollout r running BESIDE rollout r + 1.

    Rollout r accumulates into its own zeroed row of a preallocated ``[capacity, RDV_NSTATS]`` matrix (``begin()``
    returns the row).  ``end()`` -- called once the rollout's launches are enqueued -- issues the row's all-reduce in
    place with ``async_op=True``: NCCL runs it on its own stream as soon as the rollout has finished, while the host
    has already enqueued the next rollout, which writes another row.  Nothing on the compute stream waits for a
    collective until ``finish()`` (or until the matrix is full and gets folded into ``total``), and no bookkeeping
    kernel (zeroing, snapshot, running sum) sits between two rollout launches.  Let the rollout leave one SM free
    (``BatchedRendezvousEnv.sm_reserve = 1``) so that the collective's kernel does not queue behind a launch that
    fills every SM.  Without a process group it degrades to a local sum."""

    def __init__(self, device, group=None, capacity: int = 1024):
        self.group = group
        self.rows = torch.zeros((int(capacity), N.NSTATS), dtype=torch.float64, device=device)
        self.total = torch.zeros(N.NSTATS, dtype=torch.float64, device=device)
        self._last = None                   # handle of the most recent collective (NCCL executes them in issue order)
        self._r = 0

    def begin(self) -> torch.Tensor:
        if self._r == self.rows.shape[0]:
            self._fold()
        return self.rows[self._r]

    def end(self):
        work = all_reduce_stats(self.rows[self._r], group=self.group, async_op=True)
        if work is not None:
            self._last = work
        self._r += 1

    def _fold(self):
        if self._last is not None:
            self._last.wait()               # the compute stream waits for the last collective, hence for all of them
            self._last = None
        if self._r:
            self.total += self.rows[:self._r].sum(dim=0)
            self.rows[:self._r].zero_()
            self._r = 0

    def finish(self) -> torch.Tensor:
        """Wait for the outstanding collectives; returns the running total over all rollouts and ranks."""
        self._fold()
        return self.total

    def reset(self):
        self._fold()
        self.total.zero_()


def stats_to_dict(stats: torch.Tensor) -> dict:
    v = stats.detach().cpu().tolist()
    d = dict(zip(N.STAT_NAMES, v))
    ep = max(d["episodes"], 1.0)
    d["mean_return"] = d["return_sum"] / ep
    d["mean_length"] = d["length_sum"] / ep
    d["success_rate"] = d["succeeded"] / ep
    d["collision_rate"] = d["collided"] / ep
    return d
