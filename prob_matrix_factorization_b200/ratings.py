"""Device-resident rating lists: the engine's replacement for ``_build_index_lists``.

The reference groups observation indices per user and per item with a Python loop
(poisson_mf_cavi.py:73-84 and its copies in the other models).  Here the (u, i, rating) list
is uploaded once and grouped on the GPU in both orientations (``by_user`` = CSR, ``by_item`` =
CSC) by libpmf_b200's stable radix sort; each orientation is an opaque ``pmf_csr`` handle.

Two refinements that the reference (single process, tiny data) has no counterpart for:

* **tiles** -- when the factor table a pass gathers from is much larger than the L2, the pass's ratings are
  split by the id range of that table into tiles of about ``tile_bytes`` (one ``pmf_csr`` per tile); the pass
  visits one tile after the other, so every gathered row comes out of the L2 and the table streams through
  DRAM once per sweep instead of once per rating (``pmf_gamma_pass_acc`` carries the row sums across tiles);
* **shards** -- on several GPUs each rank holds the ratings of ONE user range (nnz-balanced): the user pass
  needs no exchange at all, the item pass leaves per-rank partial row sums that are added across ranks
  (``pmf_gamma_combine``).  Every rank uploads 1/world of the list; the ratings are routed to their owners with
  one all-to-all, so set-up time and memory per rank are O(nnz / world).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _cabi

DEFAULT_SEG_LEN = None  # None = auto_seg_len(nnz)


def auto_seg_len(nnz):
    """Longest run of observations one lane group handles alone, for ONE launch over `nnz` observations.

    Rows longer than this are cut into segments whose partial sums are combined by a second kernel.
    Long segments are fine for throughput (segments are scheduled longest first), but a lane group walks its
    segment 8 observations per memory round trip (~1 us under load), so the longest segment is the critical path
    of the launch: 1024 observations ~ 0.1 ms.  nnz/32768 clamped to [64, 1024] keeps that below ~10 % of the launch
    (26 ps per observation) and leaves enough independent segments for ~150 SMs x 64 groups.  Tiled / chunked
    passes are several launches: the rule is applied to the observations of one launch (measured at 2 GPUs, 16 launches
    per item pass: 3.05 ms with the whole shard's 1024, profiles/README.md).
    """
    want = (int(nnz) // 32768 + 7) // 8 * 8
    return max(64, min(1024, want))


def as_id_array(a, name):
    """Host ids -> contiguous int32 NumPy (reference: ``to_numpy(dtype=int)``)."""
    a = np.asarray(a)
    if a.dtype == np.int64:
        from .host_draws import ids_to_int32        # large arrays: cast + range check by all host cores
        fast = ids_to_int32(a, name)
        if fast is not None:
            return fast
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
    """One orientation of (a part of) the rating list (wraps a ``pmf_csr*``)."""

    def __init__(self, handle, device):
        self._h = handle
        self.device = device

    @classmethod
    def build(cls, key, other, val, n_rows, seg_len=DEFAULT_SEG_LEN, row_offset=0):
        """Group by ``key`` (ids already rebased to [0, n_rows)); ``row_offset`` = global id of local row 0."""
        if seg_len is None:
            seg_len = auto_seg_len(key.numel())
        assert key.is_cuda and key.dtype == torch.int32 and other.dtype == torch.int32 and val.dtype == torch.float32
        out = C.c_void_p()
        with torch.cuda.device(key.device):
            _cabi.call("pmf_csr_build", key.data_ptr(), other.data_ptr(), val.data_ptr(), key.numel(), n_rows,
                       seg_len, _cabi.stream_ptr(), C.byref(out))
            if row_offset:
                _cabi.call("pmf_csr_set_row_offset", out, int(row_offset))
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

    def workspace_bytes(self, ld):
        return max(int(_cabi.load().pmf_gamma_pass_workspace_bytes(self.handle, ld)), 16)

    def workspace(self, ld):
        return torch.empty(self.workspace_bytes(ld) // 4, dtype=torch.float32, device=self.device)


DEFAULT_TILE_MB = 128   # target size of the slice of a factor table one tile gathers from (about the 126 MB L2; measured best at C5)


def tile_bounds(lo, hi, row_bytes, tile_bytes=None, n_tiles=None):
    """Equal-row split of the id range [lo, hi) into tiles of about ``tile_bytes`` of factor rows (host; int64[T+1]).

    One tile when the range is at most 1.5 tiles large (tiling costs a round trip of the running row sums per
    extra tile, which only pays off once most gathers would otherwise miss the L2)."""
    rows = max(0, int(hi) - int(lo))
    if n_tiles is None:
        if tile_bytes is None:
            tile_bytes = int(os.environ.get("PMF_TILE_MB", DEFAULT_TILE_MB)) << 20
        total = rows * int(row_bytes or 0)
        n_tiles = 1 if total <= 1.5 * tile_bytes else -(-total // tile_bytes)
    n_tiles = int(max(1, min(n_tiles, 256, max(rows, 1))))
    return np.array([int(lo) + rows * t // n_tiles for t in range(n_tiles + 1)], dtype=np.int64)


def coo_partition(u_d, i_d, x_d, bounds, by_item):
    """Stable split of device triples by the bucket of u (or i): returns (u, i, x) reordered and int64 offsets."""
    n = u_d.numel()
    nb = len(bounds) - 1
    b32 = np.ascontiguousarray(bounds, dtype=np.int32)
    offs = np.zeros(nb + 1, dtype=np.int64)
    uo, io, xo = torch.empty_like(u_d), torch.empty_like(i_d), torch.empty_like(x_d)
    with torch.cuda.device(u_d.device):
        _cabi.call("pmf_coo_partition", u_d.data_ptr(), i_d.data_ptr(), x_d.data_ptr(), n, int(bool(by_item)),
                   b32.ctypes.data_as(_cabi.c_i32p), nb, uo.data_ptr(), io.data_ptr(), xo.data_ptr(),
                   offs.ctypes.data_as(C.POINTER(C.c_int64)), _cabi.stream_ptr())
    return uo, io, xo, offs


class DeviceRatings:
    """(u, i, rating) uploaded and grouped by user and by item, optionally tiled and sharded (module docstring).

    ``user_tiles[t]`` : ratings of this rank's users whose ITEM lies in item tile t, grouped by user (rows = the rank's
                        user range, ``row_offset`` = its first user; columns = global item ids)  -> user pass
    ``item_tiles[t]`` : ratings of this rank's users whose USER lies in user tile t, grouped by item (rows = all items;
                        columns = global user ids)                                               -> item pass
    ``by_user`` / ``by_item`` are the single tiles of an untiled list (what the Gaussian / extended models use).
    ``shard=(rank, world)``: this rank keeps the ratings of users [user_bounds[rank], user_bounds[rank+1]).
    ``shard_input``: "full" = every rank passes the same complete list (it uploads only its 1/world slice),
    "chunk" = every rank passes its own consecutive piece of the list (rank order = list order).
    ``row_bytes`` = bytes of one factor row (enables tiling; None = never tile).
    """

    def __init__(self, u, i, x, n_users, n_items, device=None, seg_len=DEFAULT_SEG_LEN, shard=None, row_bytes=None,
                 tile_bytes=None, user_pass_tiles=None, item_pass_tiles=None, shard_input="full", item_chunks=None):
        _cabi.require_cuda()
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.device = device
        self.n_users, self.n_items = int(n_users), int(n_items)
        self.rank, self.world = (0, 1) if shard is None else (int(shard[0]), int(shard[1]))
        u_h = u if isinstance(u, torch.Tensor) else as_id_array(u, "user")
        i_h = i if isinstance(i, torch.Tensor) else as_id_array(i, "item")
        if isinstance(x, torch.Tensor):
            x_h = x
        else:
            from .host_draws import to_float32
            x_h = to_float32(x)
        if not (len(u_h) == len(i_h) == len(x_h)):
            raise ValueError("u, i, rating must have equal length")
        if self.world > 1:
            u_d, i_d, x_d = self._route_to_owner(u_h, i_h, x_h, shard_input)
        else:
            u_d = to_device(u_h, device, torch.int32)
            i_d = to_device(i_h, device, torch.int32)
            x_d = to_device(x_h, device, torch.float32)
            self.nnz = u_d.numel()
            self.h2d_bytes = self.nnz * 12
            self.user_bounds = np.array([0, self.n_users], dtype=np.int64)
        self.nnz_local = u_d.numel()
        self.user_lo, self.user_hi = int(self.user_bounds[self.rank]), int(self.user_bounds[self.rank + 1])
        # several GPUs: the item rows can be processed in chunks so that the cross-rank combine of one chunk overlaps the
        # pass over the next (GammaEngine.item_pass); the number is fixed here because it sets the size of a launch.
        # Default 1: measured on 8 B200 (profiles/README.md) chunking never paid -- pass and combine are both latency-bound,
        # so running them side by side only shares the SMs, and each chunk adds a kernel tail + a barrier.
        if item_chunks is None:
            item_chunks = int(os.environ.get("PMF_ITEM_CHUNKS", 1)) if self.world > 1 else 1
        self.item_chunks = max(1, min(int(item_chunks), self.n_items)) if self.world > 1 else 1
        n_own = self.user_hi - self.user_lo
        env = lambda k: int(os.environ[k]) if os.environ.get(k) else None
        upt = user_pass_tiles if user_pass_tiles is not None else env("PMF_USER_PASS_TILES")
        if self.world > 1 and upt is None:
            # several GPUs: the user pass is tiled over the item chunks, so that tile c of the NEXT sweep's user pass only
            # needs the combine of chunk c to be complete (the combines of the later chunks overlap it)
            from .parallel import item_chunk_bounds
            self.item_tile_bounds = np.asarray(item_chunk_bounds(self.n_items, self.item_chunks), dtype=np.int64)
        else:
            self.item_tile_bounds = tile_bounds(0, self.n_items, row_bytes, tile_bytes, upt)
        self.user_tile_bounds = tile_bounds(self.user_lo, self.user_hi, row_bytes, tile_bytes,
                                            item_pass_tiles if item_pass_tiles is not None else env("PMF_ITEM_PASS_TILES"))
        launches_u = len(self.item_tile_bounds) - 1
        launches_i = (len(self.user_tile_bounds) - 1) * self.item_chunks
        self.seg_len_user = seg_len if seg_len is not None else auto_seg_len(self.nnz_local // launches_u)
        self.seg_len_item = seg_len if seg_len is not None else auto_seg_len(self.nnz_local // launches_i)
        self.seg_len = self.seg_len_user
        from ._engine import Trace
        tr = Trace()
        with torch.cuda.device(device):
            u_loc = u_d - self.user_lo if self.user_lo else u_d     # rows of the user pass are local to the rank's range
            self.user_tiles = self._build_tiles(u_loc, i_d, x_d, self.item_tile_bounds, by_item=True, key_is_user=True,
                                                n_rows=n_own, row_offset=self.user_lo, seg_len=self.seg_len_user) if n_own > 0 else []
            del u_loc                                               # columns of the item pass are global user ids
            tr.mark("  build: user-pass lists")
            self.item_tiles = self._build_tiles(u_d, i_d, x_d, self.user_tile_bounds, by_item=False, key_is_user=False,
                                                n_rows=self.n_items, row_offset=0, seg_len=self.seg_len_item)
            tr.mark("  build: item-pass lists")
        del u_d, i_d, x_d

    # -- construction helpers -------------------------------------------------------------------
    def _build_tiles(self, u_d, i_d, x_d, bounds, by_item, key_is_user, n_rows, row_offset, seg_len):
        """One Grouped per tile of `bounds` (ranges of the OTHER side's ids); the key side is grouped."""
        T = len(bounds) - 1
        if T > 1:
            u_p, i_p, x_p, offs = coo_partition(u_d, i_d, x_d, bounds, by_item)
        else:
            u_p, i_p, x_p, offs = u_d, i_d, x_d, np.array([0, u_d.numel()], dtype=np.int64)
        tiles = []
        for t in range(T):
            a, b = int(offs[t]), int(offs[t + 1])
            key, other = (u_p[a:b], i_p[a:b]) if key_is_user else (i_p[a:b], u_p[a:b])
            tiles.append(Grouped.build(key, other, x_p[a:b], n_rows, seg_len, row_offset))
        return tiles

    def _route_to_owner(self, u_h, i_h, x_h, shard_input):
        """Upload this rank's piece of the list and exchange ratings so that each rank holds its user range's."""
        import torch.distributed as dist
        from .parallel import balanced_bounds_from_counts
        from ._engine import Trace
        tr = Trace()
        dev, W, r = self.device, self.world, self.rank
        if shard_input == "full":
            n = len(u_h)
            per = (n + W - 1) // W
            lo, hi = min(r * per, n), min((r + 1) * per, n)
            u_h, i_h, x_h = u_h[lo:hi], i_h[lo:hi], x_h[lo:hi]
        elif shard_input != "chunk":
            raise ValueError("shard_input must be 'full' or 'chunk'")
        u_c = to_device(u_h, dev, torch.int32)
        i_c = to_device(i_h, dev, torch.int32)
        x_c = to_device(x_h, dev, torch.float32)
        self.h2d_bytes = u_c.numel() * 12
        tr.mark("  route: upload of this rank's piece")
        counts = torch.empty(self.n_users, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _cabi.call("pmf_count_keys", u_c.data_ptr(), u_c.numel(), self.n_users, counts.data_ptr(), _cabi.stream_ptr())
        dist.all_reduce(counts)                                        # ratings per user over the whole list
        self.user_bounds = balanced_bounds_from_counts(counts.cpu().numpy(), W)
        tot = torch.tensor([u_c.numel()], dtype=torch.int64, device=dev)
        dist.all_reduce(tot)
        self.nnz = int(tot.item())
        tr.mark("  route: ratings per user (count + all-reduce) and owner ranges")
        # stable split of the piece by owner, then one all-to-all: pieces arrive in rank order = list order, so every
        # rank ends with its users' ratings in ORIGINAL order (what the reference's per-row lists contain)
        u_s, i_s, x_s, offs = coo_partition(u_c, i_c, x_c, self.user_bounds, by_item=False)
        send = torch.from_numpy(np.diff(offs)).to(dev)
        recv = torch.empty(W, dtype=torch.int64, device=dev)
        dist.all_to_all_single(recv, send)
        send_l, recv_l = [int(v) for v in np.diff(offs)], [int(v) for v in recv.cpu().tolist()]
        tr.mark("  route: stable split by owner + counts exchange")
        n_loc = sum(recv_l)
        out = []
        for src in (u_s, i_s, x_s):
            dst = torch.empty(n_loc, dtype=src.dtype, device=dev)
            dist.all_to_all_single(dst, src, recv_l, send_l)
            out.append(dst)
        tr.mark("  route: all-to-all of (u, i, rating)")
        return out

    # -- views ---------------------------------------------------------------------------------------
    def _single(self, tiles, what):
        if len(tiles) != 1:
            raise RuntimeError(f"{what} is split into {len(tiles)} tiles; use the tile list")
        return tiles[0]

    by_user = property(lambda self: self._single(self.user_tiles, "by_user") if self.user_tiles else None)
    by_item = property(lambda self: self._single(self.item_tiles, "by_item"))

    def all_lists(self):
        return list(self.user_tiles) + list(self.item_tiles)

    def free(self):
        for g in self.all_lists():
            g.free()
