"""Host glue with the semantics of cbir/utils.py: the progress loop that *extends* its result
list (utils.py:8-27) and the image key = sha1 of the pixel buffer (utils.py:30-36)."""
from __future__ import annotations

import datetime
import hashlib
import sys
import time

import numpy as np


def show_progress(func, iterable, quiet: bool = False, **kwargs):
    """Apply ``func`` to every item, ``extend``-ing one flat result list like the reference does
    (utils.py:23), printing an ETA line unless ``quiet``."""
    results = []
    total = len(iterable)
    t0 = time.time()
    for i, item in enumerate(iterable):
        results.extend(func(item, **kwargs))
        if not quiet:
            avg = (time.time() - t0) / (i + 1)
            eta = datetime.timedelta(seconds=avg * (total - i - 1))
            sys.stdout.write("Progress %d/%d - ETA: %s\r" % (i + 1, total, eta))
    return results


def get_image_id(array) -> str:
    """sha1 hex digest of the array's buffer (C-contiguous arrays, like the reference)."""
    return hashlib.sha1(np.ascontiguousarray(array)).hexdigest()


def is_cuda_capable() -> bool:
    """True when a CUDA device this build can run on is visible (cbir/utils.py:39-50 asks the same
    question for its AlexNet encoder).  The kernels here are sm_100a only."""
    import torch
    if not torch.cuda.is_available():
        return False
    return all(torch.cuda.get_device_capability(d)[0] >= 10 for d in range(torch.cuda.device_count()))
