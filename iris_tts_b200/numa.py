"""Host placement for one-process-per-GPU runs: bind the process to the CPUs that are local to its GPU (NVML's affinity mask),
so the page-locked staging buffers it allocates afterwards, and the threads that fill them, live on the GPU's own NUMA node.
Without it a rank may stage its 14 MB waveform per step through the other socket.  The reference has no multi-GPU path
(src/iris/hifigan_pretrained.py:203 is a single device), so there is nothing to mirror."""
from __future__ import annotations

import os
from typing import List, Optional


def gpu_local_cpus(index: int) -> Optional[List[int]]:
    """CPU ids NVML reports as local to GPU ``index`` (None if NVML or the query is unavailable)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
            words = (os.cpu_count() + 63) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        finally:
            pynvml.nvmlShutdown()
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        return cpus or None
    except Exception:  # noqa: BLE001
        return None


def bind_process_to_gpu(index: int) -> Optional[List[int]]:
    """Restrict this process to the intersection of its current affinity and the GPU's local CPUs; returns the new CPU list, or
    None when nothing was changed (no NVML, empty intersection, or a platform without sched_setaffinity)."""
    cpus = gpu_local_cpus(index)
    if not cpus or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        allowed = sorted(set(os.sched_getaffinity(0)) & set(cpus))
        if not allowed or len(allowed) == len(os.sched_getaffinity(0)):
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except OSError:
        return None
