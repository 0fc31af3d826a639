"""Host -> device staging shared by the engines beside the main path (band noise estimator, legacy RoE, DSD emulator).

A batch arrives as a list of separate host arrays.  `upload_clips` copies them into ONE page-locked buffer that is kept
between calls (grow-only; a few threads -- numpy copies release the GIL) and sends it to the device in one asynchronous
copy.  np.concatenate into fresh pageable memory followed by a pageable copy cost ~0.3 s for 512 x 60 s of int16 PCM,
twenty times the kernels it feeds."""
from __future__ import annotations

from typing import Sequence

import numpy as np

_pins = {}      # (device index, dtype) -> page-locked torch tensor


def upload_clips(torch, dev, clips: Sequence[np.ndarray], lens: np.ndarray):
    """The clips back to back as one device tensor of their dtype (int16 or float32); at least one element long."""
    total = int(np.sum(lens))
    dt = torch.int16 if clips[0].dtype == np.int16 else torch.float32
    key = (dev.index, dt)
    pin = _pins.get(key)
    if pin is None or pin.numel() < total:
        _pins[key] = pin = torch.empty(max(total, 1), dtype=dt).pin_memory()
    host = pin.numpy()
    offs = np.concatenate(([0], np.cumsum(np.asarray(lens, dtype=np.int64))))
    n = len(clips)

    def put(rng):
        for c in rng:
            if offs[c + 1] > offs[c]:
                np.copyto(host[offs[c]:offs[c + 1]], clips[c])
    nthr = min(8, max(1, total // (4 << 20)), max(n, 1))
    if nthr <= 1:
        put(range(n))
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(nthr) as ex:
            list(ex.map(put, [range(i, n, nthr) for i in range(nthr)]))
    d = torch.empty(max(total, 1), dtype=dt, device=dev)
    if total:
        d[:total].copy_(pin[:total], non_blocking=True)
        # the staging buffer is reused by the next call: the copy must have left it before this one returns
        torch.cuda.current_stream(dev).synchronize()
    else:
        d.zero_()
    return d
