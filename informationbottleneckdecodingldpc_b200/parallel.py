"""Multi-GPU plumbing: one process per GPU, frames sharded across ranks, NCCL (over
NVLink/NVSwitch) used only to all-reduce the error counters so that every rank takes the same
``while errors < min_errors`` decision as the single-device reference loop
(Regular_LDPC_Decoding/BPSK/BER_simulation_OpenCL.py:98).  Codewords are independent: there is
no collective inside a decode.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_distributed(n_gpus: int = 1, backend: str | None = None):
    """Join the torchrun rendezvous if there is one.  Returns (rank, world_size, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device(f"cuda:{local}"))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(local_rank: int, sysfs: str = "/sys") -> int | None:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (sysfs: the PCI device's
    numa_node and that node's cpulist), so that the pinned host buffers of the host-buffer path
    (first touch) and the copy-issuing thread are local to the GPU's root complex.  With 8 ranks on a
    two-socket host the default placement puts half of the DMA traffic across the socket link.
    Returns the node, or None when the topology is not exposed (then nothing is changed)."""
    try:
        props = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open(os.path.join(sysfs, "bus/pci/devices", bdf, "numa_node")) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(os.path.join(sysfs, "devices/system/node", f"node{node}", "cpulist")) as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def shard_frames(total_frames: int, rank: int, world: int):
    """Contiguous frame range [lo, hi) of this rank (remainder spread over the first ranks)."""
    base, rem = divmod(int(total_frames), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_counters(counters, world: int | None = None):
    """Sum a short list of integer counters {bit_errors, frame_errors, frames, iterations} over all
    ranks (int64 payload of a few bytes: latency-bound, issue it once per batch, never per iteration).
    Returns a list of Python ints.  No-op without a process group."""
    vals = [int(c) for c in counters]
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return vals
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor(vals, dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(x) for x in t.tolist()]


# ---- the same all-reduce through the C ABI (ibldpc_nccl_*: NCCL bound at run time by libibldpc.so) ----------------
def nccl_library_path() -> str | None:
    """Path of the NCCL shared library PyTorch ships (None if it cannot be located): lets libibldpc.so dlopen the
    very copy the process already uses when the bare soname does not resolve."""
    import glob
    base = os.path.dirname(os.path.dirname(torch.__file__))
    for pat in ("nvidia/nccl/lib/libnccl.so.2", "torch/lib/libnccl.so.2"):
        hits = glob.glob(os.path.join(base, pat))
        if hits:
            return hits[0]
    return None


def init_counter_allreduce(decoder) -> bool:
    """Create the library-side NCCL communicator of ``decoder``'s handle (``ibldpc_nccl_init``): rank 0 draws the
    unique id, torch.distributed broadcasts its 128 bytes.  Returns False when there is no process group."""
    import ctypes as C
    from . import _lib
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return False
    if "IBLDPC_NCCL_LIB" not in os.environ:
        p = nccl_library_path()
        if p:
            os.environ["IBLDPC_NCCL_LIB"] = p
    L = _lib.lib()
    rank, world = dist.get_rank(), dist.get_world_size()
    ident = (C.c_uint8 * 128)()
    if rank == 0:
        _lib.check(L.ibldpc_nccl_unique_id(ident))
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor(list(ident), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0)
    ident = (C.c_uint8 * 128)(*t.cpu().tolist())
    _lib.check(L.ibldpc_nccl_init(decoder._ensure_handle(), ident, rank, world))
    decoder._nccl_ready = True
    return True


def counter_allreduce_fn(decoder):
    """In-place sum of an int64 CUDA counter tensor over all ranks, on the current stream.  Uses the library's own
    communicator (C ABI) when IBLDPC_ABI_ALLREDUCE=1 or the decoder's communicator is already initialised, else
    torch.distributed (the plumbing default)."""
    import ctypes as C
    from . import _lib
    use_abi = getattr(decoder, "_nccl_ready", False) or os.environ.get("IBLDPC_ABI_ALLREDUCE") == "1"
    if use_abi and not getattr(decoder, "_nccl_ready", False):
        use_abi = init_counter_allreduce(decoder)
    if not use_abi:
        return lambda c: dist.all_reduce(c, op=dist.ReduceOp.SUM)
    L = _lib.lib()

    def fn(c):
        _lib.check(L.ibldpc_allreduce_counters(decoder._ensure_handle(), C.c_void_p(c.data_ptr()), int(c.numel()),
                                               C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return fn
