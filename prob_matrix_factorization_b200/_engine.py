"""Device-side state and sweep loops behind the drop-in model classes.

Python here is host orchestration only: every numeric step on the training path is a
libpmf_b200 kernel launched through ctypes on torch-owned device memory.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
import time

import numpy as np
import torch

from . import _cabi
from .parallel import RowExchange, item_chunk_bounds, owned_item_ranges
from .ratings import DeviceRatings, Grouped, as_id_array, to_device

MAX_DEVICE_LABELS = 64
PMF_ACC_IN, PMF_ACC_OUT = 1, 2     # include/pmf_b200.h


class Trace:
    """PMF_TRACE=1: wall-clock of the phases of a fit on rank 0 (synchronises; diagnostics only)."""

    on = bool(os.environ.get("PMF_TRACE"))

    def __init__(self):
        self.t = time.perf_counter()

    def mark(self, what):
        if Trace.on:
            torch.cuda.synchronize()
            now = time.perf_counter()
            if int(os.environ.get("RANK", 0)) == 0:
                print(f"[pmf trace] {what}: {(now - self.t) * 1e3:.1f} ms", file=sys.stderr, flush=True)
            self.t = now


def row_stride(K):
    return _cabi.load().pmf_row_stride(int(K))


def pad_table(host, ld, device):
    """(R, K) host array (float64 or float32) -> float32 (R, ld) CUDA tensor, zero padded.

    The H2D copy is issued straight from the caller's buffer (asynchronous when that memory is
    pinned); the cast and the padding run on the device.
    """
    host = np.asarray(host)
    R, K = host.shape
    if host.dtype == np.float64 and host.size >= (1 << 20):
        from .host_draws import to_float32          # halve the H2D bytes; the cast runs on all host cores
        host = to_float32(np.ascontiguousarray(host))
    src = torch.from_numpy(np.ascontiguousarray(host))
    if src.dtype == torch.float32 and K == ld:
        return src.to(device, non_blocking=True)
    dev = src.to(device, non_blocking=True)
    out = torch.zeros((R, ld), dtype=torch.float32, device=device)
    out[:, :K] = dev
    return out


def table_to_host(t, K, dtype=np.float64):
    return t[:, :K].to(torch.float64 if dtype == np.float64 else torch.float32).cpu().numpy()


def normalise_ids(ids, n):
    """Reference predict() uses NumPy fancy indexing: negative ids wrap around once."""
    ids = np.asarray(ids)
    if ids.dtype.kind not in "iu":
        ids = ids.astype(np.int64)
    ids = ids.astype(np.int64, copy=True)
    neg = ids < 0
    if neg.any():
        ids[neg] += n
        if (ids < 0).any():
            raise IndexError("index out of bounds")
    return np.minimum(ids, np.iinfo(np.int32).max).astype(np.int32)


class EvalSet:
    """A (u, i, rating) frame staged on the device for repeated evaluation."""

    def __init__(self, users, items, y, n_users, n_items, device, drop_invalid=False, user_range=None):
        """``user_range=(lo, hi, is_last)`` (multi-GPU fit in progress): keep only the rows whose user this rank owns
        -- unseen users (id >= n_users) go to the last rank -- and mark the set ``sharded``: its statistics are summed
        over the ranks, so every rank sees the same numbers (and takes the same early-stopping decision)."""
        users = np.asarray(users)
        items = np.asarray(items)
        y = np.asarray(y, dtype=np.float64)
        self.drop_invalid = bool(drop_invalid)
        self.sharded = user_range is not None
        u32, i32 = normalise_ids(users, n_users), normalise_ids(items, n_items)
        labels_all = None
        if self.sharded:
            # labels are those of the WHOLE frame on every rank (of its rows with seen ids when those are dropped)
            labels_all = np.unique(y[(u32 < n_users) & (i32 < n_items)] if drop_invalid else y)
            lo, hi, is_last = user_range
            uc = np.minimum(u32, n_users)
            keep = (uc >= lo) & ((uc < hi) | bool(is_last))
            u32, i32, y = u32[keep], i32[keep], y[keep]
        self.n = len(y)
        y_for_labels = y
        if drop_invalid:  # gaussian_mf_cavi_bias.py:323-324 filters before np.unique sees the labels
            ok = (u32 < n_users) & (i32 < n_items)
            y_for_labels = y[ok]
        labels = np.unique(y_for_labels) if labels_all is None else labels_all
        self.labels = labels
        self.on_device = 0 < len(labels) <= MAX_DEVICE_LABELS
        self.y_host = y
        self.u = to_device(u32, device)
        self.i = to_device(i32, device)
        self.y = to_device(y.astype(np.float32), device)
        if self.on_device:
            lab = np.searchsorted(labels, y).astype(np.int32)
            lab[(lab >= len(labels)) | (labels[np.minimum(lab, len(labels) - 1)] != y)] = -1
            self.label = to_device(lab, device)
            self.n_labels = len(labels)
        else:
            self.label, self.n_labels = None, 0
        self.out = torch.zeros(4 + 2 * MAX_DEVICE_LABELS, dtype=torch.float64, device=device)
        self.scratch = torch.zeros((int(_cabi.load().pmf_eval_stats_scratch_bytes()) + 7) // 8, dtype=torch.int64, device=device)


def eval_stats_launch(ev, F_user, F_item, n_users, n_items, K, ld, b_user=None, b_item=None, global_mean=0.0):
    """Enqueue the fused evaluation kernel of one EvalSet (statistics land in ``ev.out``; no host round trip)."""
    with torch.cuda.device(F_user.device):
        _cabi.call("pmf_eval_stats", ev.u.data_ptr(), ev.i.data_ptr(), ev.y.data_ptr(), _cabi.ptr(ev.label),
                   ev.n_labels, ev.n, F_user.data_ptr(), n_users, F_item.data_ptr(), n_items, K, ld,
                   _cabi.ptr(b_user), _cabi.ptr(b_item), float(global_mean), int(ev.drop_invalid),
                   ev.out.data_ptr(), ev.scratch.data_ptr(), _cabi.stream_ptr())


_SIDE_STREAMS = {}


def side_stream(dev, which="loop", high_priority=False):
    """One long-lived non-default stream per (device, purpose): creating a stream costs 25-900 ms on a cold context
    (measured inside fits, profiles/README.md), far more than a small fit.  ``high_priority``: the block scheduler
    hands freed SM slots to this stream's kernels first -- what lets the cross-rank combine of one item chunk run
    WHILE the pass over the next chunk (whose grid fills the GPU many times over) is still executing."""
    dev = torch.device(dev)
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), which)
    if key not in _SIDE_STREAMS:
        prio = torch.cuda.Stream.priority_range()[1] if high_priority else 0
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev, priority=prio)
    return _SIDE_STREAMS[key]


class DeviceLoop:
    """The fit loop as one CUDA graph with a device-side WHILE (pmf_loop_*, csrc/loop.cu): sweeps, validation statistics
    and the reference's early-stopping rule run back to back on the GPU; the host reads the iteration count and the
    RMSE history once, at the end.

        loop = DeviceLoop(dev, max_iter, ev_out=ev.out, rule=0, tol=cfg.tol)
        with loop.body():            # ONE iteration is enqueued (captured, not executed)
            engine.sweep(...); eval_stats_launch(ev, ...)
        n_iter, rmse_history = loop.run()
    """

    def __init__(self, dev, max_iter, ev_out=None, rule=0, tol=None):
        self.dev = torch.device(dev)
        self.max_iter = int(max_iter)
        self.ev_out, self.rule, self.tol = ev_out, int(rule), tol
        self.stream = side_stream(self.dev, "loop")
        self.iter = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.history = torch.zeros(max(1, self.max_iter), dtype=torch.float64, device=self.dev)
        self._h = C.c_void_p()

    def body(self):
        import contextlib

        @contextlib.contextmanager
        def ctx():
            self.stream.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.device(self.dev), torch.cuda.stream(self.stream):
                _cabi.call("pmf_loop_begin", self.stream.cuda_stream, C.byref(self._h))
                try:
                    yield self
                    _cabi.call("pmf_loop_decide", self._h, _cabi.ptr(self.ev_out), self.rule,
                               0.0 if self.tol is None else float(self.tol), int(self.tol is not None), self.max_iter,
                               self.iter.data_ptr(), self.history.data_ptr(), self.stream.cuda_stream)
                    _cabi.call("pmf_loop_end", self._h)
                except BaseException:
                    self.free()
                    raise
        return ctx()

    def run(self):
        """Launch the loop and wait for it: (iterations executed, validation RMSE per iteration as float64 NumPy)."""
        with torch.cuda.device(self.dev):
            self.iter.zero_()
            self.stream.wait_stream(torch.cuda.current_stream(self.dev))
            _cabi.call("pmf_loop_run", self._h, self.stream.cuda_stream)
            torch.cuda.current_stream(self.dev).wait_stream(self.stream)
            n = int(self.iter.cpu().item())
            hist = self.history[:n].cpu().numpy() if self.ev_out is not None else np.zeros(0)
        return n, hist

    def free(self):
        if self._h is not None and self._h.value:
            _cabi.load().pmf_loop_free(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def device_loop_enabled():
    return os.environ.get("PMF_DEVICE_LOOP", "1") != "0"


def eval_stats(ev, F_user, F_item, n_users, n_items, K, ld, b_user=None, b_item=None, global_mean=0.0):
    """RMSE / macro-MAE / MAE / Poisson LPL of one EvalSet: one fused kernel + one tiny D2H."""
    eval_stats_launch(ev, F_user, F_item, n_users, n_items, K, ld, b_user, b_item, global_mean)
    with torch.cuda.device(F_user.device):
        if ev.sharded:
            import torch.distributed as dist
            dist.all_reduce(ev.out)                          # float64 sums over the ranks' rows: identical everywhere
        out = ev.out.cpu().numpy()
    cnt = out[0]
    res = {"count": cnt, "sse": float(out[1]), "rmse": float(np.sqrt(out[1] / cnt)) if cnt > 0 else float("nan"),
           "mae": float(out[2] / cnt) if cnt > 0 else float("nan"), "poisson_lpl": float(out[3])}
    if ev.on_device:
        sae, c = out[4:4 + ev.n_labels], out[4 + ev.n_labels:4 + 2 * ev.n_labels]
        res["macro_mae"] = float(np.mean(sae[c > 0] / c[c > 0])) if (c > 0).any() else float("nan")
    elif ev.sharded:
        raise NotImplementedError(f"more than {MAX_DEVICE_LABELS} distinct ratings in a sharded evaluation")
    else:
        # more distinct true values than the fused kernel tracks: reduce predictions per label on host
        pred = predict(ev.u, ev.i, F_user, F_item, n_users, n_items, K, ld, b_user, b_item, global_mean)
        y = ev.y_host + global_mean
        if ev.drop_invalid:
            ok = ((ev.u < n_users) & (ev.i < n_items)).cpu().numpy()
            y, pred = y[ok], pred[ok]
        per = [np.mean(np.abs(y[y == lab] - pred[y == lab])) for lab in np.unique(y)]
        res["macro_mae"] = float(np.mean(per)) if per else float("nan")
    return res


def predict(u_dev, i_dev, F_user, F_item, n_users, n_items, K, ld, b_user=None, b_item=None, global_mean=0.0,
            softplus=False):
    """Device ids -> float64 NumPy predictions (pmf_predict)."""
    n = u_dev.numel()
    out = torch.empty(n, dtype=torch.float64, device=F_user.device)
    with torch.cuda.device(F_user.device):
        _cabi.call("pmf_predict", u_dev.data_ptr(), i_dev.data_ptr(), n, F_user.data_ptr(), n_users,
                   F_item.data_ptr(), n_items, K, ld, _cabi.ptr(b_user), _cabi.ptr(b_item), float(global_mean),
                   int(bool(softplus)), out.data_ptr(), _cabi.stream_ptr())
    return out.cpu().numpy()


# Symmetric-memory tables (multi-GPU) are kept for the life of the process and handed to the next engine of the same
# shape: allocating + rendezvous of the two tables costs ~0.5 s per fit and releasing them another ~0.4 s (measured at
# 2 GPUs, profiles/README.md) -- more than 20 sweeps of the 100 M-rating config.  One entry per table name: a fit
# with other shapes replaces it.
_SYMM_CACHE = {}


class GammaEngine:
    """Poisson MF / HPF-CAVI state on one GPU (one user-range shard of the ratings) and its sweep.

    Sweep order follows SURVEY.md Appendix A: user pass (old E_theta, E_beta[, E_xi]) -> E_theta
    [-> xi] -> item pass (NEW E_theta, old E_beta[, E_eta]) -> E_beta [-> eta]; two dependent
    SDDMM+reduce passes per iteration.  Each pass visits its tiles (ratings.py) in turn; on several
    GPUs the item pass ends with the cross-rank combine of the row sums:

      "mc"   pmf_gamma_combine: multimem.ld_reduce adds the ranks' sums inside the NVSwitch, the owner updates the row
             and multimem.st replicates it (default beyond 2 GPUs);
      "ce"   pmf_gamma_combine_staged: the sums are copied to their owner by the copy engines and added in rank order
             (default at 2 GPUs); update + replication as "mc";
      "nccl" unfused baseline: NCCL all-reduce of the sums, every rank updates every row (fall-back without multicast).
    The item rows can be processed in `item_chunks` chunks, the exchange of one chunk (side stream) next to the pass over
    the next -- measured not to pay with the present kernels (default 1 chunk).
    """

    def __init__(self, ratings: DeviceRatings, K, user_shape, item_shape, user_rate=None, item_rate=None,
                 hyper=None, keep_params=True, exchange=None, item_chunks=None):
        self.r = ratings
        self.dev = ratings.device
        self.K = int(K)
        self.ld = row_stride(K)
        self.N, self.M = ratings.n_users, ratings.n_items
        self.world, self.rank = ratings.world, ratings.rank
        self.user_lo, self.user_hi = ratings.user_lo, ratings.user_hi
        self.user_shape, self.item_shape = float(user_shape), float(item_shape)
        self.user_rate = None if user_rate is None else float(user_rate)
        self.item_rate = None if item_rate is None else float(item_rate)
        # hyper = dict(user_shape=a_xi, user_rate_prior=b', item_shape=a_eta, item_rate_prior=d') for HPF
        self.hyper = hyper
        self.keep_params = keep_params
        f = lambda rows, cols=self.ld: torch.zeros((rows, cols), dtype=torch.float32, device=self.dev)
        if exchange is None:
            # measured (profiles/README.md): 2 GPUs -- staging by the copy engines 2.94 ms vs in-switch 3.08 ms per sweep;
            # 8 GPUs -- in-switch 1.16 ms vs staged 1.90 ms (seven small peer copies per rank reach only ~200 GB/s)
            exchange = os.environ.get("PMF_EXCHANGE") or ("ce" if ratings.world == 2 else "mc")
        if exchange not in ("mc", "ce", "nccl"):
            raise ValueError("exchange must be 'mc', 'ce' or 'nccl'")
        self.exchange = exchange if self.world > 1 else "none"
        self._symm = {}
        self._side = None
        self._ready = None        # events of an item pass whose combines are still in flight (see item_pass(join=False))
        self.staged = self.exchange == "ce"      # "ce": sums staged on their owner by the copy engines (else in-switch reduce)
        if self.exchange == "ce":
            self.exchange = "mc"                 # same protocol (barriers, chunks, multicast replication of the new rows)
        if self.exchange == "mc":
            import torch.distributed as dist
            ok, why = 1.0, ""
            try:
                self._setup_multicast()
            except Exception as exc:  # multicast needs NVSwitch + driver support
                ok, why = 0.0, str(exc)
            flag = torch.tensor([ok], device=self.dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)          # the fallback decision is collective
            if flag.item() < 1.0:
                import warnings
                warnings.warn(f"multicast combine unavailable ({why or 'on another rank'}); using the NCCL all-reduce")
                for entry in self._symm.values():
                    entry[3]["busy"] = False
                self._symm = {}
                self.exchange = "nccl"
        # E_theta: during a sharded fit only this rank's rows [user_lo, user_hi) are live (nobody else reads them)
        self.E_theta = f(self.N)
        n_item_tiles = len(ratings.item_tiles)
        if self.exchange == "mc":
            self.E_beta, self.acc_item = self._symm["E_beta"][0], self._symm["acc_item"][0]
            self._side = side_stream(self.dev, "combine", high_priority=True)
            self._copy_streams = [side_stream(self.dev, f"copy{k}") for k in range(3)] if self.staged else []
        else:
            self.E_beta = f(self.M)
            self.acc_item = f(self.M, 2 * self.ld) if (self.world > 1 or n_item_tiles > 1) else None
        self.acc_user = f(self.user_hi - self.user_lo, 2 * self.ld) if len(ratings.user_tiles) > 1 else None
        self.shp_theta = f(self.N) if keep_params else None
        self.rte_theta = f(self.N) if keep_params else None
        self.shp_beta = f(self.M) if keep_params else None
        self.rte_beta = f(self.M) if keep_params else None
        if hyper is not None:
            v = lambda rows: torch.zeros(rows, dtype=torch.float32, device=self.dev)
            self.rate_xi, self.rate_eta = v(self.N), v(self.M)
            self.E_xi, self.E_eta = v(self.N), v(self.M)
        else:
            self.rate_xi = self.E_xi = self.rate_eta = self.E_eta = None
        # item rows in chunks (multicast combine only): chunk c's combine overlaps chunk c+1's pass
        if item_chunks is None:
            item_chunks = ratings.item_chunks
        self.item_chunks = max(1, min(int(item_chunks), self.M)) if self.exchange == "mc" else 1
        if self.item_chunks > 1:
            cb = item_chunk_bounds(self.M, self.item_chunks)
            self.item_lists = [[g.slice(cb[c], cb[c + 1]) for g in ratings.item_tiles] for c in range(self.item_chunks)]
        else:
            self.item_lists = [list(ratings.item_tiles)]
        self.owned_items = owned_item_ranges(self.M, self.item_chunks, self.world, self.rank) if self.world > 1 else []
        ws_bytes = lambda lists: max([g.workspace_bytes(self.ld) for g in lists] + [16])
        w = lambda nbytes: torch.empty(nbytes // 4, dtype=torch.float32, device=self.dev)
        self.ws_user = w(ws_bytes(ratings.user_tiles))
        self.ws_item = w(ws_bytes([g for lst in self.item_lists for g in lst]))
        self.xu = RowExchange(ratings.user_bounds) if self.world > 1 else None
        self._item_params_synced = True
        count = lambda g: (1 if g.n_segments > 0 else 0) + (1 if g.n_multi_rows > 0 else 0)
        self.launches_per_sweep = (sum(count(g) for g in ratings.user_tiles)
                                   + sum(count(g) for lst in self.item_lists for g in lst)
                                   + (self.item_chunks if self.world > 1 else 0))

    # -- state upload --------------------------------------------------------------------------
    def load_means(self, E_theta, E_beta, E_xi=None, E_eta=None):
        """Upload the initial expectations (host float64, drawn by NumPy exactly as the reference)."""
        if self.world > 1:
            # this rank's user rows only; E_beta is replicated: 1/world over PCIe each, the rest over NVLink
            from .parallel import replicate_from_slices
            lo, hi = self.user_lo, self.user_hi
            if hi > lo:
                self.E_theta[lo:hi].copy_(pad_table(np.asarray(E_theta)[lo:hi], self.ld, self.dev))
            full = replicate_from_slices(np.asarray(E_beta), self.dev, self.world, self.rank)
            if full.shape[1] == self.ld and full.dtype == torch.float32:
                self.E_beta.copy_(full)
            else:
                self.E_beta.zero_()
                self.E_beta[:, :full.shape[1]] = full
        else:
            self.E_theta.copy_(pad_table(E_theta, self.ld, self.dev))
            self.E_beta.copy_(pad_table(E_beta, self.ld, self.dev))
        if self.hyper is not None:
            self.E_xi.copy_(to_device(np.asarray(E_xi, dtype=np.float32), self.dev))
            self.E_eta.copy_(to_device(np.asarray(E_eta, dtype=np.float32), self.dev))
        if self.world > 1:
            torch.cuda.current_stream(self.dev).synchronize()
            import torch.distributed as dist
            dist.barrier()                    # every replica of E_beta is in place before any rank's first pass

    def download_means(self, out_theta, out_beta, owned_only=False):
        """Copy E_theta / E_beta (first K columns) into caller-provided (pinned) float32 host tensors.

        ``owned_only`` (multi-GPU): copy just this rank's user rows and its 1/world share of the (replicated) item
        rows -- together the ranks' copies are the whole result.  Returns the number of bytes copied."""
        K = self.K
        nbytes = 0
        item_share = (self.M * self.rank // self.world, self.M * (self.rank + 1) // self.world)
        for out, tab, own in ((out_theta, self.E_theta, (self.user_lo, self.user_hi)), (out_beta, self.E_beta, item_share)):
            lo, hi = own if owned_only else (0, tab.shape[0])
            src = tab[lo:hi] if self.ld == K else tab[lo:hi, :K]
            out[lo:hi].copy_(src, non_blocking=True)
            nbytes += (hi - lo) * K * 4
        torch.cuda.current_stream(self.dev).synchronize()
        return nbytes

    def eval_range(self):
        """``user_range`` for EvalSet while a sharded fit is running (None on one GPU)."""
        if self.world == 1 or self.exchange == "closed":
            return None
        return (self.user_lo, self.user_hi, self.rank == self.world - 1)

    # -- one pass ------------------------------------------------------------------------------
    def _pass(self, grouped, E_oth, E_self, shp, rte, shape_prior, rate_prior, rate_vec, hyper_rate, hyper_mean,
              hyper_shape, hyper_rate_prior, ws, acc=None, acc_base=0, flags=0):
        _cabi.call("pmf_gamma_pass_acc", grouped.handle, self.K, self.ld, E_oth.data_ptr(), E_self.data_ptr(),
                   _cabi.ptr(shp), _cabi.ptr(rte), shape_prior, 0.0 if rate_prior is None else rate_prior,
                   _cabi.ptr(rate_vec), _cabi.ptr(hyper_rate), _cabi.ptr(hyper_mean), hyper_shape,
                   hyper_rate_prior, _cabi.ptr(ws), _cabi.ptr(acc) if flags else None, acc_base, flags,
                   _cabi.stream_ptr())

    def _stage_copies(self, c, done):
        """Chunk c: copy the sums of the rows each peer owns into that peer's staging table (copy engines; no SMs)."""
        stage, hdl = self._symm["stage"][0], self._symm["stage"][1]
        if not hasattr(self, "_peer_stage"):
            self._peer_stage = {p: hdl.get_buffer(p, tuple(stage.shape), torch.float32) for p in range(self.world) if p != self.rank}
            C_ = self.item_chunks
            self._owned_all = {p: owned_item_ranges(self.M, C_, self.world, p) for p in range(self.world)}
            self._row0_all = {p: np.concatenate([[0], np.cumsum([hi - lo for lo, hi in self._owned_all[p]])]) for p in range(self.world)}
        row_bytes = 2 * self.ld * 4
        events = []
        for k, cs in enumerate(self._copy_streams):
            cs.wait_event(done)
            for p in range(self.world):
                if p == self.rank or (p % len(self._copy_streams)) != k:
                    continue
                lo, hi = self._owned_all[p][c]
                if hi > lo:
                    dst = self._peer_stage[p][self.rank, int(self._row0_all[p][c])]
                    _cabi.call("pmf_memcpy_async", dst.data_ptr(), self.acc_item[lo].data_ptr(), (hi - lo) * row_bytes, cs.cuda_stream)
            ev = torch.cuda.Event()
            ev.record(cs)
            events.append(ev)
        return events

    def _combine(self, lo, hi, write_params, multicast, chunk=0):
        h = self.hyper
        wp = write_params and self.keep_params
        if multicast and self.staged:
            stage = self._symm["stage"][0]
            _cabi.call("pmf_gamma_combine_staged", lo, hi, self.K, self.ld, self.acc_item.data_ptr(), 0, stage.data_ptr(),
                       self.world, self._stage_rows * 2 * self.ld, self.rank, int(self._row0_all[self.rank][chunk]),
                       self.E_beta.data_ptr(), self._symm["E_beta"][2],
                       _cabi.ptr(self.shp_beta if wp else None), _cabi.ptr(self.rte_beta if wp else None),
                       self.item_shape, 0.0 if self.item_rate is None else self.item_rate, _cabi.ptr(self.E_eta),
                       _cabi.ptr(self.rate_eta), _cabi.ptr(self.E_eta), h["item_shape"] if h else 0.0,
                       h["item_rate_prior"] if h else 0.0, _cabi.stream_ptr())
            return
        _cabi.call("pmf_gamma_combine", lo, hi, self.K, self.ld, self.acc_item.data_ptr(),
                   self._symm["acc_item"][2] if multicast else None, 0, self.E_beta.data_ptr(),
                   self._symm["E_beta"][2] if multicast else None,
                   _cabi.ptr(self.shp_beta if wp else None), _cabi.ptr(self.rte_beta if wp else None),
                   self.item_shape, 0.0 if self.item_rate is None else self.item_rate, _cabi.ptr(self.E_eta),
                   _cabi.ptr(self.rate_eta), _cabi.ptr(self.E_eta), h["item_shape"] if h else 0.0,
                   h["item_rate_prior"] if h else 0.0, _cabi.stream_ptr())

    def _setup_multicast(self):
        """E_beta and the item-side row sums in torch symmetric memory with an NVSwitch multicast alias (plumbing)."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        group = dist.group.WORLD
        tables = [("E_beta", (self.M, self.ld)), ("acc_item", (self.M, 2 * self.ld))]
        if self.staged:
            # staging table: slot s = the sums rank s computed for the rows THIS rank owns (owner's rows in chunk order)
            self._stage_rows = -(-self.M // self.world) + 64       # a rank owns M/world rows, +-1 per chunk
            tables.append(("stage", (self.world, self._stage_rows, 2 * self.ld)))
        for name, shape in tables:
            key = (name, self.dev.index)
            hit = _SYMM_CACHE.get(key)
            if hit is not None and not hit["busy"] and tuple(hit["t"].shape) == shape and hit["group"] == group.group_name:
                hit["busy"] = True
                self._symm[name] = (hit["t"], hit["hdl"], hit["mc"], hit)
                continue
            t = symm_mem.empty(shape, dtype=torch.float32, device=self.dev)
            hdl = symm_mem.rendezvous(t, group.group_name)
            if not hdl.multicast_ptr and name != "stage":
                raise RuntimeError("no multicast pointer")
            t.zero_()
            entry = {"t": t, "hdl": hdl, "mc": int(hdl.multicast_ptr or 0), "group": group.group_name, "busy": True}
            if hit is None or not hit["busy"]:
                _SYMM_CACHE[key] = entry          # (an entry in use by a live engine is left alone; this one is not cached)
            self._symm[name] = (t, hdl, entry["mc"], entry)
        torch.cuda.synchronize(self.dev)

    def _rank_barrier(self):
        """Device-side signal-pad barrier on the current stream: the ranks' work enqueued before it is complete after it."""
        self._symm["E_beta"][1].barrier(channel=0)

    def user_pass(self, write_params=True):
        h = self.hyper
        tiles = self.r.user_tiles
        wp = write_params and self.keep_params
        main = torch.cuda.current_stream(self.dev)
        # combines of the previous item pass that are still in flight: tile t gathers E_beta rows of item chunk t only
        per_tile = self._ready if (self._ready and len(self._ready) == len(tiles)) else None
        if self._ready and per_tile is None:
            self._join()
        for t, g in enumerate(tiles):
            if per_tile is not None:
                main.wait_event(per_tile[t])
            last = t == len(tiles) - 1
            flags = (PMF_ACC_IN if t > 0 else 0) | (0 if last else PMF_ACC_OUT)
            self._pass(g, self.E_beta, self.E_theta, self.shp_theta if wp else None, self.rte_theta if wp else None,
                       self.user_shape, self.user_rate, self.E_xi, self.rate_xi, self.E_xi,
                       h["user_shape"] if h else 0.0, h["user_rate_prior"] if h else 0.0, self.ws_user,
                       self.acc_user, self.user_lo, flags)
        if per_tile is not None:
            self._join()

    def _join(self):
        """The current stream waits for the cross-rank combines still running on the side stream."""
        if self._ready:
            torch.cuda.current_stream(self.dev).wait_stream(self._side)
            self._ready = None

    def item_pass(self, write_params=True, join=True):
        """``join=False`` (multicast combine): return with the combines still in flight on the side stream; the next
        ``user_pass`` waits for them tile by tile (every other consumer of E_beta must call ``_join`` first)."""
        h = self.hyper
        wp = write_params and self.keep_params
        self._item_params_synced = self.exchange != "mc"
        main = torch.cuda.current_stream(self.dev)
        self._join()
        ready = []
        for c, tiles in enumerate(self.item_lists):
            for t, g in enumerate(tiles):
                last = t == len(tiles) - 1 and self.world == 1
                flags = (PMF_ACC_IN if t > 0 else 0) | (0 if last else PMF_ACC_OUT)
                self._pass(g, self.E_theta, self.E_beta, self.shp_beta if wp else None, self.rte_beta if wp else None,
                           self.item_shape, self.item_rate, self.E_eta, self.rate_eta, self.E_eta,
                           h["item_shape"] if h else 0.0, h["item_rate_prior"] if h else 0.0, self.ws_item,
                           self.acc_item, 0, flags)
            if self.exchange == "mc":
                done = torch.cuda.Event()
                done.record(main)
                if self.staged:
                    for ev in self._stage_copies(c, done):
                        self._side.wait_event(ev)
                self._side.wait_event(done)
                with torch.cuda.stream(self._side):
                    # one barrier says: every rank has parked its sums of chunk c AND finished its combine of chunk c-1
                    self._rank_barrier()
                    if c > 0:
                        ready.append(torch.cuda.Event())
                        ready[-1].record(self._side)            # E_beta rows of chunk c-1 are final on every replica
                    lo, hi = self.owned_items[c]
                    self._combine(lo, hi, write_params, multicast=True, chunk=c)
        if self.exchange == "mc":
            with torch.cuda.stream(self._side):
                self._rank_barrier()                            # every rank's new E_beta rows have landed everywhere
                ready.append(torch.cuda.Event())
                ready[-1].record(self._side)
            self._ready = ready
            if join:
                self._join()
        elif self.exchange == "nccl":
            import torch.distributed as dist
            dist.all_reduce(self.acc_item)
            self._combine(0, self.M, write_params, multicast=False)

    def sweep(self, write_params=True):
        """One CAVI iteration.  ``write_params=False`` skips materialising the Gamma shape/rate tables (only the
        last sweep's are observable; 1.3 GB of writes per sweep at the 100 M-rating config)."""
        with torch.cuda.device(self.dev):
            self.user_pass(write_params)
            self.item_pass(write_params)

    def sweeps(self, n, write_params_last=True):
        """``n`` iterations back to back.  On several GPUs the sweeps are software-pipelined: the combines of an item pass
        run on the side stream while the next sweep's user pass works through the item chunks that are already final."""
        with torch.cuda.device(self.dev):
            for s in range(n):
                wp = write_params_last and s == n - 1
                self.user_pass(wp)
                self.item_pass(wp, join=s == n - 1)

    def close(self):
        """Detach from the symmetric-memory tables (multi-GPU): state becomes ordinary device tensors and the tables go
        back to the cache for the next engine.  Every rank calls it at the same point of its program; no barrier is
        needed: the last item pass ended with one (all remote stores into this replica have landed), the copy below is
        stream-ordered, and the next engine's first remote store comes after a barrier that orders it behind this copy."""
        if self._symm:
            self.E_beta = self._symm["E_beta"][0].clone()
            self.acc_item = None
            for entry in self._symm.values():
                entry[3]["busy"] = False
            self._symm = {}
        if self.exchange != "none":
            self.exchange = "closed"

    def _sync_item_params(self):
        """Multicast combine: shape/rate/hyper rows of an item are written by its owner only -> make them complete
        everywhere (the non-owned rows are zeroed and the tables summed over the ranks)."""
        if self._item_params_synced or self.exchange != "mc":
            return
        import torch.distributed as dist
        mask = torch.zeros(self.M, dtype=torch.bool, device=self.dev)
        for lo, hi in self.owned_items:
            mask[lo:hi] = True
        tabs = [t for t in ((self.rate_eta, self.E_eta) if self.hyper else ()) if t is not None]
        if self.keep_params:
            tabs += [self.shp_beta, self.rte_beta]
        for t in tabs:
            t[~mask] = 0
            dist.all_reduce(t)
        self._item_params_synced = True

    def sync_params(self):
        """Multi-GPU: make every table complete on every rank (end of fit: E_theta rows and the Gamma parameters live
        with their owners only while the sweeps run)."""
        if self.xu is None or self.exchange == "closed":
            return
        self.xu.gather(self.E_theta)
        if self.hyper:
            self.xu.gather(self.rate_xi, self.E_xi)
        if self.keep_params:
            self.xu.gather(self.shp_theta, self.rte_theta)
        self._sync_item_params()

    # -- a11 extras (parity unpinned) ----------------------------------------------------------------
    def load_params(self, shp_theta, rte_theta, shp_beta, rte_beta, rate_xi, rate_eta):
        """Upload Gamma shape/rate tables (needed before the first sweep by the digamma pass and the ELBO)."""
        for name, host in (("shp_theta", shp_theta), ("rte_theta", rte_theta), ("shp_beta", shp_beta), ("rte_beta", rte_beta)):
            getattr(self, name).copy_(pad_table(host, self.ld, self.dev))
        self.rate_xi.copy_(to_device(np.asarray(rate_xi, dtype=np.float32), self.dev))
        self.rate_eta.copy_(to_device(np.asarray(rate_eta, dtype=np.float32), self.dev))

    def geomean_tables(self):
        """G = exp(psi(shape))/rate for both sides (pmf_gamma_geomean)."""
        if getattr(self, "G_theta", None) is None:
            self.G_theta, self.G_beta = torch.zeros_like(self.E_theta), torch.zeros_like(self.E_beta)
        with torch.cuda.device(self.dev):
            for shp, rte, G, rows in ((self.shp_theta, self.rte_theta, self.G_theta, self.N),
                                      (self.shp_beta, self.rte_beta, self.G_beta, self.M)):
                _cabi.call("pmf_gamma_geomean", shp.data_ptr(), rte.data_ptr(), rows, self.K, self.ld, G.data_ptr(),
                           _cabi.stream_ptr())
        return self.G_theta, self.G_beta

    def sweep_digamma(self):
        """One sweep with the multinomial (digamma) allocation of docs/Models.tex:652-664 (single GPU, untiled)."""
        if self.world > 1:
            raise NotImplementedError("digamma allocation is single-GPU")
        h = self.hyper
        with torch.cuda.device(self.dev):
            for grouped, G_oth, E_oth, G_self, E_self, shp, rte, shape, rate_vec, hr, hs, hp, ws in (
                    (self.r.by_user, self.G_beta, self.E_beta, self.G_theta, self.E_theta, self.shp_theta, self.rte_theta,
                     self.user_shape, self.E_xi, self.rate_xi, h["user_shape"], h["user_rate_prior"], self.ws_user),
                    (self.r.by_item, self.G_theta, self.E_theta, self.G_beta, self.E_beta, self.shp_beta, self.rte_beta,
                     self.item_shape, self.E_eta, self.rate_eta, h["item_shape"], h["item_rate_prior"], self.ws_item)):
                _cabi.call("pmf_gamma_pass_digamma", grouped.handle, self.K, self.ld, G_oth.data_ptr(), E_oth.data_ptr(),
                           G_self.data_ptr(), E_self.data_ptr(), shp.data_ptr(), rte.data_ptr(), shape, 0.0,
                           rate_vec.data_ptr(), hr.data_ptr(), rate_vec.data_ptr(), hs, hp, _cabi.ptr(ws), _cabi.stream_ptr())

    def elbo(self, cfg, refresh_geomean=True):
        """Observed-only HPF ELBO and its six components (pmf_hpf_elbo); one D2H of 6 doubles.

        Several GPUs: each rank adds the likelihood of its own ratings, the prior/entropy terms of its own users and of
        a 1/world share of the items; the six sums are all-reduced (item-side parameters are completed first)."""
        sharded = self.world > 1 and self.exchange != "closed"
        if sharded:
            self._sync_item_params()
        if refresh_geomean or getattr(self, "G_theta", None) is None:
            self.geomean_tables()
        tiles = self.r.user_tiles
        if not tiles:       # a rank without users still owes the row terms of its share of the items
            z = torch.zeros(1, dtype=torch.int32, device=self.dev)
            self._empty_list = Grouped.build(z[:0], z[:0], z[:0].float(), 1, 8)
            tiles = [self._empty_list]
        out = torch.zeros((len(tiles), 6), dtype=torch.float64, device=self.dev)
        u_rng = (self.user_lo, self.user_hi) if sharded else (0, self.N)
        i_rng = (self.M * self.rank // self.world, self.M * (self.rank + 1) // self.world) if sharded else (0, self.M)
        with torch.cuda.device(self.dev):
            for t, g in enumerate(tiles):
                ur, ir = (u_rng, i_rng) if t == 0 else ((0, 0), (0, 0))     # row terms once, likelihood per tile
                _cabi.call("pmf_hpf_elbo", g.handle, self.K, self.ld, self.E_theta.data_ptr(), self.E_beta.data_ptr(),
                           self.G_theta.data_ptr(), self.G_beta.data_ptr(), self.shp_theta.data_ptr(), self.rte_theta.data_ptr(),
                           self.shp_beta.data_ptr(), self.rte_beta.data_ptr(), self.rate_xi.data_ptr(), self.rate_eta.data_ptr(),
                           ur[0], ur[1], ir[0], ir[1], cfg.a, cfg.a_prime, cfg.b_prime, cfg.c, cfg.c_prime, cfg.d_prime,
                           out[t].data_ptr(), _cabi.stream_ptr())
        parts_d = out.sum(0)
        if sharded:
            import torch.distributed as dist
            dist.all_reduce(parts_d)
        parts = parts_d.cpu().numpy()
        return float(parts.sum()), parts

    # -- accounting ----------------------------------------------------------------------------
    def algorithmic_bytes_per_sweep(self):
        """per pass nnz*(4K+8) + R*(16K+4) (+12 R for the HPF hyper vectors)."""
        K, nnz = self.K, self.r.nnz
        per_row = 16 * K + 4 + (12 if self.hyper is not None else 0)
        return 2 * nnz * (4 * K + 8) + (self.N + self.M) * per_row
