"""N>1 host path on CPU: world_size-2 gloo processes shard a clip list and all-gather the
per-clip statistics rows (the path's only exchange step)."""
import os
import socket

import numpy as np
import pytest

from audio_processing_tools_b200.parallel import shard_by_samples, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, counts, q):
    import torch
    import torch.distributed as dist
    from audio_processing_tools_b200.parallel import gather_clip_stats
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = counts[rank]
    local = torch.zeros((n, 8), dtype=torch.float32)
    local[:, 0] = torch.arange(n)                    # plan-local clip index
    local[:, 1] = 100 * rank + torch.arange(n)       # "rain_frame_count"
    out = gather_clip_stats(local, counts)
    q.put((rank, out.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("counts", [(3, 3), (4, 2)])
def test_gather_clip_stats_world2(counts):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, list(counts), q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(got[0], got[1])
    assert np.array_equal(got[0][:, 0], np.arange(sum(counts)))          # global clip ids, in order
    expect = np.concatenate([100 * r + np.arange(c) for r, c in enumerate(counts)])
    assert np.array_equal(got[0][:, 1], expect)


def test_shard_helpers():
    assert [shard_range(1000, r, 8) for r in range(8)] == [(125 * r, 125 * (r + 1)) for r in range(8)]
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 2), (2, 5), (5, 7), (7, 10)]
    lengths = [10, 10, 10, 10, 40, 40]
    parts = shard_by_samples(lengths, 2)
    assert parts[0][0] == 0 and parts[-1][1] == len(lengths) and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    tot = [sum(lengths[a:b]) for a, b in parts]
    assert max(tot) <= 80


def test_bind_near_gpu_is_harmless_without_a_gpu():
    """No NVML device here: the call must leave the affinity alone and say so."""
    import os
    from audio_processing_tools_b200.parallel import bind_near_gpu
    before = os.sched_getaffinity(0)
    assert bind_near_gpu(0) is None
    assert os.sched_getaffinity(0) == before
