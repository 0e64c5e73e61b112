"""Worker threads for host-side work that releases the GIL (Arrow dictionary decoding, numpy column arithmetic, staging
copies of pageable input arrays).  One small pool per process, none on a single-core host."""
from __future__ import annotations

import os

_POOL = []


def cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def pool():
    """The shared ``ThreadPoolExecutor``, or None when the process may only use one core."""
    n = cores()
    if n < 2:
        return None
    if not _POOL:
        from concurrent.futures import ThreadPoolExecutor
        _POOL.append(ThreadPoolExecutor(max_workers=min(12, n), thread_name_prefix="laris-host"))
    return _POOL[0]
