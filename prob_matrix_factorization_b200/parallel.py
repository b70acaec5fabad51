"""Multi-GPU plumbing: one process per GPU, ratings sharded by nonzero, factors replicated.

The reference is single-process; this is new.  Each rank owns a
contiguous, nnz-balanced range of user rows (for the user pass) and of item rows (for the item
pass).  After a pass every rank holds fresh factor rows only for its own range, so the per-pass
"statistics combine" is an all-gather of owned rows of the replicated table (variable sizes),
not an all-reduce of zero-padded statistics: 2.5x fewer bytes over NVLink for the same result
(SURVEY.md §8e).  ``torch.distributed`` is the transport (NCCL on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist


def dist_info():
    """(rank, world, local_rank) from torchrun's environment; (0, 1, 0) when not launched by it."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0)))


def init_process_group(backend=None):
    """Join the default process group if WORLD_SIZE > 1 (idempotent)."""
    rank, world, local = dist_info()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def balanced_row_bounds(row_ptr, parts):
    """Host mirror of pmf_csr_partition: row-aligned, nnz-balanced boundaries (int64[parts+1]).

    Boundary p is the first row whose start offset is >= nnz*p/parts, never before boundary p-1.
    """
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    n_rows, nnz = len(row_ptr) - 1, int(row_ptr[-1])
    bounds = np.zeros(parts + 1, dtype=np.int64)
    for p in range(1, parts):
        target = (nnz * p) // parts
        lo = int(np.searchsorted(row_ptr[:n_rows], target, side="left"))
        bounds[p] = max(lo, bounds[p - 1])
    bounds[parts] = n_rows
    return bounds


def replicate_from_slices(host, device, world, rank, dtype=None, group=None):
    """Full device copy of a host array that every rank holds: each rank uploads only its 1/world slice over PCIe
    and the slices are all-gathered over NVLink (instead of `world` full H2D copies through the same host)."""
    n = int(host.shape[0])
    per = (n + world - 1) // world
    lo, hi = min(rank * per, n), min((rank + 1) * per, n)
    src = host[lo:hi] if isinstance(host, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(host[lo:hi]))
    if dtype is not None and src.dtype != dtype:
        src = src.to(dtype)
    full = torch.empty((per * world,) + tuple(src.shape[1:]), dtype=src.dtype, device=device)
    mine = full[rank * per:(rank + 1) * per]
    mine[:hi - lo].copy_(src, non_blocking=True)
    dist.all_gather_into_tensor(full, mine, group=group)
    return full[:n]


class RowExchange:
    """All-gather of each rank's owned row range of a replicated, row-major table (in place)."""

    def __init__(self, bounds, group=None):
        self.bounds = [int(b) for b in bounds]
        self.group = group
        self.world = len(self.bounds) - 1
        self.rank = dist.get_rank(group) if (dist.is_initialized() and self.world > 1) else 0
        self.bytes_last = 0

    def owned(self, table, rank=None):
        r = self.rank if rank is None else rank
        return table[self.bounds[r]:self.bounds[r + 1]]

    def gather(self, *tables):
        """After the call every rank holds every rank's owned rows of each table."""
        if self.world == 1:
            return
        self.bytes_last = 0
        backend = dist.get_backend(self.group)
        for t in tables:
            views = [t[self.bounds[g]:self.bounds[g + 1]] for g in range(self.world)]
            self.bytes_last += sum(v.numel() * v.element_size() for g, v in enumerate(views) if g != self.rank)
            if backend == "nccl":
                # uneven sizes: ProcessGroupNCCL issues one coalesced group of broadcasts
                nonempty = all(v.numel() > 0 for v in views)
                if nonempty:
                    dist.all_gather(views, views[self.rank], group=self.group)
                    continue
            for g, v in enumerate(views):
                if v.numel() > 0:
                    dist.broadcast(v, src=dist.get_global_rank(self.group, g) if self.group else g, group=self.group)


class _CudaArray:
    """Minimal __cuda_array_interface__ carrier so torch can view library-allocated device memory."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerTable:
    """A replicated float32 table whose copies on the other ranks are mapped into this process (CUDA IPC).

    ``local`` is a torch view of this rank's copy; ``peer_ptrs`` are the other ranks' copies as raw device
    pointers reachable over NVLink (P2P loads/stores).  Used by the fused pass+exchange kernel
    (``pmf_gamma_pass_p2p``): each rank stores the rows it owns straight into every replica.
    """

    def __init__(self, shape, device, group=None):
        import ctypes as C

        from . import _cabi
        self.device = torch.device(device)
        self.shape = tuple(int(s) for s in shape)
        nbytes = 4 * int(np.prod(self.shape))
        self._ptr = C.c_void_p()
        handle = C.create_string_buffer(64)
        with torch.cuda.device(self.device):
            _cabi.call("pmf_ipc_alloc", max(nbytes, 64), C.byref(self._ptr), handle)
            self.local = torch.as_tensor(_CudaArray(self._ptr.value, self.shape), device=self.device)
            self.local.zero_()
        world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        handles = [None] * world
        dist.all_gather_object(handles, handle.raw, group=group)
        self._opened = []
        self.peer_ptrs = []
        with torch.cuda.device(self.device):
            for r, h in enumerate(handles):
                if r == self.rank:
                    continue
                p = C.c_void_p()
                _cabi.call("pmf_ipc_open", C.create_string_buffer(h, 64), C.byref(p))
                self._opened.append(p)
                self.peer_ptrs.append(p.value)
        self.peer_array = (C.c_void_p * max(1, len(self.peer_ptrs)))(*self.peer_ptrs)

    def close(self):
        from . import _cabi
        lib = _cabi.load()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            for p in self._opened:
                lib.pmf_ipc_close(p)
            self._opened = []
            if self._ptr is not None and self._ptr.value:
                self.local = None
                lib.pmf_ipc_free(self._ptr)
                self._ptr = None
