"""Device plumbing: PyTorch owns device memory and streams, nothing else.

Every helper here raises if CUDA is unavailable -- there is no CPU path.
"""
import warnings

import numpy as np
import torch

from . import _lib


_cuda_ok = False


def require_cuda():
    """Raises unless a CUDA device and the native library are there.  Checked once per process: the helpers below
    call this on every allocation, and torch.cuda.is_available() costs microseconds each time (a headline analyze
    call made ~50 of them on its launch-bound multi-GPU prologue)."""
    global _cuda_ok
    if _cuda_ok:
        return
    if not torch.cuda.is_available():
        raise RuntimeError("salib_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    _lib.lib()
    _cuda_ok = True


def device():
    require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return 0 if t is None else t.data_ptr()


def to_device(a, dtype=None, non_blocking=False):
    """numpy / torch -> contiguous device tensor (no copy if already there).  ``non_blocking``: a pinned host
    tensor is uploaded asynchronously on the current stream (the caller keeps it alive until the result is read)."""
    dev = device()
    if isinstance(a, torch.Tensor):
        if a.device != dev or (dtype and a.dtype != dtype):
            t = a.to(device=dev, dtype=dtype, non_blocking=bool(non_blocking and a.device.type == "cpu" and a.is_pinned()))
        else:
            t = a
        return t.contiguous()
    arr = np.ascontiguousarray(a)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)   # read-only host arrays (np.load views) are only read here
        t = torch.from_numpy(arr)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.to(dev, non_blocking=bool(non_blocking and t.is_pinned()))


def to_device_sharded(a, dtype, min_bytes=64 << 20):
    """Host array that every rank of the torch.distributed job holds -> full device copy on every rank,
    moved as world_size shards over PCIe (one per rank) and all-gathered over NVLink.  Falls back to
    ``to_device`` for device tensors, single-process runs and small arrays."""
    import torch.distributed as dist

    rank, world_size = world()
    if world_size == 1 or (isinstance(a, torch.Tensor) and a.is_cuda):
        return to_device(a, dtype)
    flat = a.reshape(-1) if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a).reshape(-1))
    if flat.numel() * flat.element_size() < min_bytes:
        return to_device(a, dtype)
    n = flat.numel()
    per = -(-n // world_size)
    lo, hi = min(rank * per, n), min((rank + 1) * per, n)
    part = torch.zeros((per,), dtype=dtype, device=device())
    part[: hi - lo] = flat[lo:hi].to(device=device(), dtype=dtype, non_blocking=True)
    full = torch.empty((world_size * per,), dtype=dtype, device=device())
    dist.all_gather_into_tensor(full, part)
    return full[:n]


def bind_host_to_gpu():
    """Pin the calling thread to the CPUs NVML reports as local to the current GPU (its NUMA node).  A torchrun
    worker that allocates its pinned staging buffers after this call gets host pages next to its own GPU: eight
    ranks uploading at once otherwise share whatever socket their processes happened to start on.  Best effort:
    returns the CPU list, or None when NVML or the affinity call is unavailable."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        uuid = torch.cuda.get_device_properties(torch.cuda.current_device()).uuid
        h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None


def empty(shape, dtype):
    return torch.empty(shape, dtype=dtype, device=device())


def zeros(shape, dtype):
    return torch.zeros(shape, dtype=dtype, device=device())


def world():
    """(rank, world_size) of the torch.distributed job, (0, 1) when not initialised."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(total, rank, world_size):
    """Contiguous [lo, hi) share of `total` units for `rank` (first ranks get the remainder)."""
    base, rem = divmod(int(total), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
