"""Multi-GPU plumbing: one process per GPU, ratings sharded by nonzero along USER ranges.

The reference is single-process; this is new.  Each rank owns a
contiguous, nnz-balanced range of users and holds exactly those users' ratings.  The user pass is
then local (no exchange: E_theta rows are only ever read by their owner's item pass); the item pass
leaves per-rank partial row sums over all items, which are added across ranks -- inside the NVSwitch
by ``pmf_gamma_combine`` (multimem.ld_reduce), or by an NCCL all-reduce in the unfused baseline -- and
the new E_beta rows are replicated.  ``torch.distributed`` is the transport for set-up and the
baseline (NCCL on GPUs, gloo in CPU tests).
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


def balanced_bounds_from_counts(counts, parts):
    """Row-aligned, nnz-balanced boundaries from per-row rating counts (host; int64[parts+1]).

    Same rule as pmf_csr_partition / balanced_row_bounds: boundary p is the first row whose start offset is
    >= nnz*p/parts."""
    counts = np.asarray(counts, dtype=np.int64)
    row_ptr = np.zeros(len(counts) + 1, dtype=np.int64)
    np.cumsum(counts, out=row_ptr[1:])
    return balanced_row_bounds(row_ptr, parts)


def item_chunk_bounds(n_items, chunks, shape=None):
    """Item-row chunks of the multi-GPU sweep, in processing order (int list, chunks+1 entries).

    The item pass parks its row sums chunk by chunk; the cross-rank combine of chunk c (high-priority side stream)
    overlaps the pass over the later chunks.  shape "equal": equal ranges -- with the user pass tiled over the SAME ranges
    the combines also overlap the next sweep's user-pass tiles;  "shrink": sizes proportional to chunks, chunks-1, ..., 1,
    so that the last chunk's combine -- the only one nothing overlaps when the user pass is not tiled -- is the smallest
    (40/30/20/10 % for 4 chunks).  Default: PMF_CHUNK_SHAPE or "shrink"."""
    chunks = max(1, min(int(chunks), int(n_items)))
    shape = shape or os.environ.get("PMF_CHUNK_SHAPE", "shrink")
    if shape == "equal":
        return [n_items * c // chunks for c in range(chunks + 1)]
    total = chunks * (chunks + 1) // 2
    bounds, acc = [0], 0
    for c in range(chunks):
        acc += chunks - c
        bounds.append(max(bounds[-1] + 1, n_items * acc // total) if c < chunks - 1 else n_items)
    for c in range(chunks - 1, 0, -1):          # tiny inputs: keep every chunk non-empty
        bounds[c] = min(bounds[c], bounds[c + 1] - 1)
    return bounds


def owned_item_ranges(n_items, chunks, world, rank, shape=None):
    """Item rows whose combine step `rank` performs: the rank's share of each chunk of item_chunk_bounds()
    (inside a chunk rank r owns the r-th of `world` equal parts)."""
    b = item_chunk_bounds(n_items, chunks, shape)
    out = []
    for c in range(len(b) - 1):
        lo, n = b[c], b[c + 1] - b[c]
        out.append((lo + n * rank // world, lo + n * (rank + 1) // world))
    return out


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
