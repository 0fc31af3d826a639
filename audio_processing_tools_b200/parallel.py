"""Multi-GPU plumbing: clips shard by file, one process per GPU, no data-path collective.

The only exchange step of the path is the final gather of the fixed-width per-clip statistics
rows (the reference's analogue is the ProcessPoolExecutor returning one dict row
per file, audio_processing_framework.py:282-285).  ``torch.distributed`` carries it: NCCL over
NVLink on the GPU box, gloo in the CPU tests.  32 B per clip -- latency-bound, not bandwidth-bound.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

N_STATS = 8


def shard_range(n_clips: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous clip index range [lo, hi) owned by ``rank`` (equal lengths => balanced)."""
    return (n_clips * rank) // world, (n_clips * (rank + 1)) // world


def shard_by_samples(lengths: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """Contiguous ranges balanced by total samples (ragged corpora): cut points at the clip whose
    cumulative sample count first reaches r/world of the total."""
    lengths = np.asarray(lengths, dtype=np.int64)
    cum = np.concatenate([[0], np.cumsum(lengths)])
    total = int(cum[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        i = int(np.searchsorted(cum, target, side="left"))
        i = min(max(i, cuts[-1]), len(lengths))
        cuts.append(i)
    cuts.append(len(lengths))
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def gather_clip_stats(local_stats, counts: Sequence[int], clip_id_base: int = 0):
    """All-gather of the per-clip statistics rows.

    local_stats: torch tensor [n_local, 8] (float32) on this rank's device (CUDA for NCCL, CPU for
    gloo); column 0 holds the plan-local clip index and is rebased to the global clip id here.
    counts: clips per rank (len == world size).  Returns a tensor [sum(counts), 8] ordered by
    global clip id, identical on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    assert len(counts) == world and local_stats.shape == (counts[rank], N_STATS)
    local = local_stats.clone()
    local[:, 0] += float(clip_id_base + sum(counts[:rank]))
    if world == 1:
        return local
    # ragged shards are padded to the largest one so that one fixed-size all-gather serves both cases
    cmax = max(counts)
    padded = torch.zeros((cmax, N_STATS), dtype=local.dtype, device=local.device)
    padded[:counts[rank]] = local
    out = torch.empty((world * cmax, N_STATS), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded)
    if len(set(counts)) == 1:
        return out
    return torch.cat([out[r * cmax:r * cmax + c] for r, c in enumerate(counts)], dim=0)


def bind_near_gpu(device_index: int):
    """Restrict the calling process to the CPUs NVML reports as local to a GPU and return the previous affinity set
    (None if NVML or the affinity call is unavailable).  With one process per GPU the pinned staging buffers of
    `apt_run_host_i16` are then allocated on the GPU's own NUMA node; on an 8-GPU box host-to-device copies from
    remote-socket memory otherwise halve the end-to-end rate.  Call before allocating pinned memory; undo with
    `os.sched_setaffinity(0, previous)`."""
    import os
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device_index)
        handle = None
        uuid = getattr(props, "uuid", None)
        if uuid is not None:
            try:
                handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
            except Exception:
                handle = None
        if handle is None:
            handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        previous = os.sched_getaffinity(0)
        cpus &= previous
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return previous
    except Exception:
        return None
