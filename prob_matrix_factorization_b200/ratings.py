"""Device-resident rating lists: the engine's replacement for ``_build_index_lists``.

The reference groups observation indices per user and per item with a Python loop
(poisson_mf_cavi.py:73-84 and its copies in the other models).  Here the (u, i, rating) list
is uploaded once and grouped on the GPU in both orientations (``by_user`` = CSR, ``by_item`` =
CSC) by libpmf_b200's stable radix sort; each orientation is an opaque ``pmf_csr`` handle.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi

DEFAULT_SEG_LEN = None  # None = auto_seg_len(nnz)


def auto_seg_len(nnz):
    """Longest run of observations one lane group handles alone.

    Rows longer than this are cut into segments whose partial sums are combined by a second kernel.
    Long segments are fine for throughput (segments are scheduled longest first); the cap only has to
    keep enough independent segments for ~150 SMs x 64 groups: nnz/32768 clamped to [64, 1024].
    """
    want = (int(nnz) // 32768 + 7) // 8 * 8
    return max(64, min(1024, want))


def as_id_array(a, name):
    """Host ids -> contiguous int32 NumPy (reference: ``to_numpy(dtype=int)``)."""
    a = np.asarray(a)
    if a.dtype != np.int32:
        if a.size and (a.min() < 0 or a.max() > np.iinfo(np.int32).max - 1):
            raise ValueError(f"{name} ids must lie in [0, 2^31-2]")
        a = a.astype(np.int32)
    return np.ascontiguousarray(a)


def to_device(a, device, dtype=None):
    """Host array -> CUDA tensor (async when the host memory is pinned/registered)."""
    t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.to(device, non_blocking=True)


class Grouped:
    """One orientation of the rating list (wraps a ``pmf_csr*``)."""

    def __init__(self, handle, device):
        self._h = handle
        self.device = device

    @classmethod
    def build(cls, key, other, val, n_rows, seg_len=DEFAULT_SEG_LEN):
        if seg_len is None:
            seg_len = auto_seg_len(key.numel())
        assert key.is_cuda and key.dtype == torch.int32 and other.dtype == torch.int32 and val.dtype == torch.float32
        out = C.c_void_p()
        with torch.cuda.device(key.device):
            _cabi.call("pmf_csr_build", key.data_ptr(), other.data_ptr(), val.data_ptr(), key.numel(), n_rows,
                       seg_len, _cabi.stream_ptr(), C.byref(out))
        return cls(out, key.device)

    def slice(self, row_begin, row_end):
        out = C.c_void_p()
        with torch.cuda.device(self.device):
            _cabi.call("pmf_csr_slice", self._h, row_begin, row_end, _cabi.stream_ptr(), C.byref(out))
        return Grouped(out, self.device)

    def partition(self, parts):
        """nnz-balanced row boundaries, int32[parts+1] (pmf_csr_partition)."""
        b = np.zeros(parts + 1, dtype=np.int32)
        with torch.cuda.device(self.device):
            _cabi.call("pmf_csr_partition", self._h, parts, b.ctypes.data_as(_cabi.c_i32p))
        return b

    def free(self):
        if self._h is not None and self._h.value:
            _cabi.load().pmf_csr_free(self._h)
        self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    @property
    def handle(self):
        if self._h is None:
            raise RuntimeError("rating list was freed")
        return self._h

    nnz = property(lambda self: _cabi.load().pmf_csr_nnz(self.handle))
    n_rows = property(lambda self: _cabi.load().pmf_csr_rows(self.handle))
    row_offset = property(lambda self: _cabi.load().pmf_csr_row_offset(self.handle))
    n_segments = property(lambda self: _cabi.load().pmf_csr_segments(self.handle))
    n_multi_rows = property(lambda self: _cabi.load().pmf_csr_multi_rows(self.handle))
    seg_len = property(lambda self: _cabi.load().pmf_csr_seg_len(self.handle))
    device_bytes = property(lambda self: _cabi.load().pmf_csr_device_bytes(self.handle))

    def _fetch(self, getter, count, dtype):
        out = np.empty(count, dtype=dtype)
        if count:
            with torch.cuda.device(self.device):
                _cabi.call("pmf_copy_to_host", out.ctypes.data, getattr(_cabi.load(), getter)(self.handle),
                           out.nbytes, _cabi.stream_ptr())
        return out

    def row_ptr(self):
        return self._fetch("pmf_csr_row_ptr", self.n_rows + 1, np.int32)

    def perm(self):
        return self._fetch("pmf_csr_perm", self.nnz, np.int32)

    def col(self):
        return self._fetch("pmf_csr_col", self.nnz, np.int32)

    def val(self):
        return self._fetch("pmf_csr_val", self.nnz, np.float32)

    def workspace(self, ld):
        nbytes = _cabi.load().pmf_gamma_pass_workspace_bytes(self.handle, ld)
        return torch.empty(max(nbytes, 16) // 4, dtype=torch.float32, device=self.device)


class DeviceRatings:
    """(u, i, rating) uploaded and grouped by user and by item.

    ``shard=(rank, world)`` keeps only this rank's nnz-balanced, row-aligned slice of each
    orientation (ratings sharded by nonzero); ``user_bounds`` / ``item_bounds`` hold
    the row ranges of every rank.
    """

    def __init__(self, u, i, x, n_users, n_items, device=None, seg_len=DEFAULT_SEG_LEN, shard=None):
        _cabi.require_cuda()
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.device = device
        self.n_users, self.n_items = int(n_users), int(n_items)
        self.rank, self.world = (0, 1) if shard is None else (int(shard[0]), int(shard[1]))
        u_h = u if isinstance(u, torch.Tensor) else as_id_array(u, "user")
        i_h = i if isinstance(i, torch.Tensor) else as_id_array(i, "item")
        x_h = x if isinstance(x, torch.Tensor) else np.asarray(x, dtype=np.float32)
        if self.world > 1 and not any(isinstance(t, torch.Tensor) and t.is_cuda for t in (u_h, i_h, x_h)):
            # every rank holds the same host list: upload 1/world each, assemble over NVLink
            from .parallel import replicate_from_slices
            u_d = replicate_from_slices(u_h, device, self.world, self.rank, torch.int32)
            i_d = replicate_from_slices(i_h, device, self.world, self.rank, torch.int32)
            x_d = replicate_from_slices(x_h, device, self.world, self.rank, torch.float32)
        else:
            u_d = to_device(u_h, device, torch.int32)
            i_d = to_device(i_h, device, torch.int32)
            x_d = to_device(x_h, device, torch.float32)
        if not (u_d.numel() == i_d.numel() == x_d.numel()):
            raise ValueError("u, i, rating must have equal length")
        self.nnz = u_d.numel()
        self.h2d_bytes = self.nnz * 12
        if seg_len is None:
            seg_len = auto_seg_len(self.nnz)
        with torch.cuda.device(device):
            by_user = Grouped.build(u_d, i_d, x_d, self.n_users, seg_len)
            by_item = Grouped.build(i_d, u_d, x_d, self.n_items, seg_len)
        del u_d, i_d, x_d
        self.user_bounds = by_user.partition(self.world)
        self.item_bounds = by_item.partition(self.world)
        if self.world > 1:
            r = self.rank
            ub, ib = self.user_bounds, self.item_bounds
            sliced_u = by_user.slice(int(ub[r]), int(ub[r + 1])) if ub[r + 1] > ub[r] else None
            sliced_i = by_item.slice(int(ib[r]), int(ib[r + 1])) if ib[r + 1] > ib[r] else None
            by_user.free()
            by_item.free()
            by_user, by_item = sliced_u, sliced_i
        self.by_user, self.by_item = by_user, by_item

    def free(self):
        for g in (self.by_user, self.by_item):
            if g is not None:
                g.free()
