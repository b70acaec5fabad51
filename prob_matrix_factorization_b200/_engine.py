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
from .parallel import PeerTable, RowExchange
from .ratings import DeviceRatings, as_id_array, to_device

MAX_DEVICE_LABELS = 64


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

    def __init__(self, users, items, y, n_users, n_items, device, drop_invalid=False):
        users = np.asarray(users)
        items = np.asarray(items)
        y = np.asarray(y, dtype=np.float64)
        self.n = len(y)
        self.drop_invalid = bool(drop_invalid)
        u32, i32 = normalise_ids(users, n_users), normalise_ids(items, n_items)
        y_for_labels = y
        if drop_invalid:  # gaussian_mf_cavi_bias.py:323-324 filters before np.unique sees the labels
            ok = (u32 < n_users) & (i32 < n_items)
            y_for_labels = y[ok]
        labels = np.unique(y_for_labels)
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


def eval_stats(ev, F_user, F_item, n_users, n_items, K, ld, b_user=None, b_item=None, global_mean=0.0):
    """RMSE / macro-MAE / MAE / Poisson LPL of one EvalSet: one fused kernel + one tiny D2H."""
    with torch.cuda.device(F_user.device):
        _cabi.call("pmf_eval_stats", ev.u.data_ptr(), ev.i.data_ptr(), ev.y.data_ptr(), _cabi.ptr(ev.label),
                   ev.n_labels, ev.n, F_user.data_ptr(), n_users, F_item.data_ptr(), n_items, K, ld,
                   _cabi.ptr(b_user), _cabi.ptr(b_item), float(global_mean), int(ev.drop_invalid),
                   ev.out.data_ptr(), _cabi.stream_ptr())
        out = ev.out.cpu().numpy()
    cnt = out[0]
    res = {"count": cnt, "rmse": float(np.sqrt(out[1] / cnt)) if cnt > 0 else float("nan"),
           "mae": float(out[2] / cnt) if cnt > 0 else float("nan"), "poisson_lpl": float(out[3])}
    if ev.on_device:
        sae, c = out[4:4 + ev.n_labels], out[4 + ev.n_labels:4 + 2 * ev.n_labels]
        res["macro_mae"] = float(np.mean(sae[c > 0] / c[c > 0])) if (c > 0).any() else float("nan")
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


class GammaEngine:
    """Poisson MF / HPF-CAVI state on one GPU (one shard of the ratings) and its sweep.

    Sweep order follows SURVEY.md Appendix A: user pass (old E_theta, E_beta[, E_xi]) -> E_theta
    [-> xi] -> item pass (NEW E_theta, old E_beta[, E_eta]) -> E_beta [-> eta]; two dependent
    SDDMM+reduce passes per iteration.
    """

    def __init__(self, ratings: DeviceRatings, K, user_shape, item_shape, user_rate=None, item_rate=None,
                 hyper=None, keep_params=True, exchange=None):
        self.r = ratings
        self.dev = ratings.device
        self.K = int(K)
        self.ld = row_stride(K)
        self.N, self.M = ratings.n_users, ratings.n_items
        self.user_shape, self.item_shape = float(user_shape), float(item_shape)
        self.user_rate = None if user_rate is None else float(user_rate)
        self.item_rate = None if item_rate is None else float(item_rate)
        # hyper = dict(user_shape=a_xi, user_rate_prior=b', item_shape=a_eta, item_rate_prior=d') for HPF
        self.hyper = hyper
        self.keep_params = keep_params
        f = lambda rows: torch.zeros((rows, self.ld), dtype=torch.float32, device=self.dev)
        # multi-GPU row exchange, all fused into the pass kernel except "nccl":
        #   "mc"   one multimem.st per row slice, replicated by the NVSwitch (default; falls back to "p2p")
        #   "p2p"  one store per peer into CUDA-IPC mapped replicas
        #   "nccl" all-gather of owned rows after the pass
        if exchange is None:
            exchange = os.environ.get("PMF_EXCHANGE", "mc")
        self.exchange = exchange if ratings.world > 1 else "none"
        self._peer = {}
        self._symm = {}
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
                warnings.warn(f"multicast row exchange unavailable ({why or 'on another rank'}); using P2P stores")
                self._symm = {}
                self.exchange = "p2p"
        if self.exchange == "mc":
            self.E_theta, self.E_beta = self._symm["E_theta"][0], self._symm["E_beta"][0]
        elif self.exchange == "p2p":
            import torch.distributed as dist
            # only the factor tables are replicated every pass: E_xi / E_eta are read solely by the rank
            # that owns the row (rate prior of its own rows), so they are gathered once, at the end
            for name, shape in (("E_theta", (self.N, self.ld)), ("E_beta", (self.M, self.ld))):
                self._peer[name] = PeerTable(shape, self.dev)
            self.E_theta, self.E_beta = self._peer["E_theta"].local, self._peer["E_beta"].local
            self._token = torch.zeros(1, dtype=torch.float32, device=self.dev)
            dist.barrier()
        else:
            self.E_theta, self.E_beta = f(self.N), f(self.M)
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
        self.ws_user = ratings.by_user.workspace(self.ld) if ratings.by_user is not None else None
        self.ws_item = ratings.by_item.workspace(self.ld) if ratings.by_item is not None else None
        self.xu = RowExchange(ratings.user_bounds) if ratings.world > 1 else None
        self.xi_ = RowExchange(ratings.item_bounds) if ratings.world > 1 else None
        self.n_peers = ratings.world - 1 if self.exchange == "p2p" else (-1 if self.exchange == "mc" else 0)
        self.launches_per_sweep = sum(
            (1 if g.n_segments > 0 else 0) + (1 if g.n_multi_rows > 0 else 0)
            for g in (ratings.by_user, ratings.by_item) if g is not None)

    # -- state upload --------------------------------------------------------------------------
    def load_means(self, E_theta, E_beta, E_xi=None, E_eta=None):
        """Upload the initial expectations (host float64, drawn by NumPy exactly as the reference)."""
        if self.r.world > 1:
            # identical host arrays on every rank: 1/world over PCIe each, the rest over NVLink
            from .parallel import replicate_from_slices
            for dst, host in ((self.E_theta, E_theta), (self.E_beta, E_beta)):
                full = replicate_from_slices(np.asarray(host), self.dev, self.r.world, self.r.rank)
                if full.shape[1] == self.ld and full.dtype == torch.float32:
                    dst.copy_(full)
                else:
                    dst.zero_()
                    dst[:, :full.shape[1]] = full
        else:
            self.E_theta.copy_(pad_table(E_theta, self.ld, self.dev))
            self.E_beta.copy_(pad_table(E_beta, self.ld, self.dev))
        if self.hyper is not None:
            self.E_xi.copy_(to_device(np.asarray(E_xi, dtype=np.float32), self.dev))
            self.E_eta.copy_(to_device(np.asarray(E_eta, dtype=np.float32), self.dev))

    def download_means(self, out_theta, out_beta, owned_only=False):
        """Copy E_theta / E_beta (first K columns) into caller-provided (pinned) float32 host tensors.

        ``owned_only`` (multi-GPU): copy just this rank's row ranges -- the replicas are identical, so the ranks'
        owned rows together are the whole result.  Returns the number of bytes copied."""
        K = self.K
        nbytes = 0
        for out, tab, bounds in ((out_theta, self.E_theta, self.r.user_bounds), (out_beta, self.E_beta, self.r.item_bounds)):
            lo, hi = (int(bounds[self.r.rank]), int(bounds[self.r.rank + 1])) if owned_only else (0, tab.shape[0])
            src = tab[lo:hi] if self.ld == K else tab[lo:hi, :K]
            out[lo:hi].copy_(src, non_blocking=True)
            nbytes += (hi - lo) * K * 4
        torch.cuda.current_stream(self.dev).synchronize()
        return nbytes

    # -- one pass ------------------------------------------------------------------------------
    def _pass(self, grouped, E_oth, E_self, shp, rte, shape_prior, rate_prior, rate_vec, hyper_rate, hyper_mean,
              hyper_shape, hyper_rate_prior, ws, peer_E=None):
        if grouped is None:
            return
        _cabi.call("pmf_gamma_pass_p2p", grouped.handle, self.K, self.ld, E_oth.data_ptr(), E_self.data_ptr(),
                   _cabi.ptr(shp), _cabi.ptr(rte), shape_prior, 0.0 if rate_prior is None else rate_prior,
                   _cabi.ptr(rate_vec), _cabi.ptr(hyper_rate), _cabi.ptr(hyper_mean), hyper_shape,
                   hyper_rate_prior, _cabi.ptr(ws), self.n_peers if peer_E is not None else 0,
                   peer_E if peer_E is not None else None, None, _cabi.stream_ptr())

    def _setup_multicast(self):
        """E_theta / E_beta in torch symmetric memory with an NVSwitch multicast alias (plumbing only)."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        group = dist.group.WORLD
        for name, rows in (("E_theta", self.N), ("E_beta", self.M)):
            t = symm_mem.empty((rows, self.ld), dtype=torch.float32, device=self.dev)
            hdl = symm_mem.rendezvous(t, group.group_name)
            if not hdl.multicast_ptr:
                raise RuntimeError("no multicast pointer")
            t.zero_()
            self._symm[name] = (t, hdl, (C.c_void_p * 1)(hdl.multicast_ptr))
        torch.cuda.synchronize(self.dev)

    def _rank_barrier(self):
        """All ranks' pass kernels (and their remote stores into this replica) are complete after this."""
        if self.exchange == "mc":
            self._symm["E_theta"][1].barrier(channel=0)      # device-side signal-pad barrier on the current stream
            return
        import torch.distributed as dist
        dist.all_reduce(self._token)

    def user_pass(self):
        h = self.hyper
        self._pass(self.r.by_user, self.E_beta, self.E_theta, self.shp_theta, self.rte_theta, self.user_shape,
                   self.user_rate, self.E_xi, self.rate_xi, self.E_xi,
                   h["user_shape"] if h else 0.0, h["user_rate_prior"] if h else 0.0, self.ws_user,
                   self._remote("E_theta"))
        if self.exchange in ("p2p", "mc"):
            self._rank_barrier()
        elif self.xu is not None:
            self.xu.gather(self.E_theta)

    def item_pass(self):
        h = self.hyper
        self._pass(self.r.by_item, self.E_theta, self.E_beta, self.shp_beta, self.rte_beta, self.item_shape,
                   self.item_rate, self.E_eta, self.rate_eta, self.E_eta,
                   h["item_shape"] if h else 0.0, h["item_rate_prior"] if h else 0.0, self.ws_item,
                   self._remote("E_beta"))
        if self.exchange in ("p2p", "mc"):
            self._rank_barrier()
        elif self.xi_ is not None:
            self.xi_.gather(self.E_beta)

    def sweep(self):
        with torch.cuda.device(self.dev):
            self.user_pass()
            self.item_pass()

    def _remote(self, name):
        """ctypes array of remote aliases of a table for pmf_gamma_pass_p2p (peers, or the multicast address)."""
        if self.exchange == "p2p":
            return self._peer[name].peer_array
        if self.exchange == "mc":
            return self._symm[name][2]
        return None

    def close(self):
        """Unmap / free peer-shared tables (multi-GPU p2p exchange).  Call on every rank."""
        if self._symm:
            import torch.distributed as dist
            torch.cuda.synchronize(self.dev)
            dist.barrier()
            for k, (t, hdl, arr) in self._symm.items():
                setattr(self, k, t.clone())
            self._symm = {}
            self.exchange = "closed"
            self.n_peers = 0
        if self._peer:
            import torch.distributed as dist
            keep = {k: getattr(self, k).clone() for k in self._peer}     # state stays readable after close
            torch.cuda.synchronize(self.dev)
            dist.barrier()
            for t in self._peer.values():
                t.close()
            self._peer = {}
            for k, v in keep.items():
                setattr(self, k, v)
            self.exchange = "closed"
            self.n_peers = 0

    def sync_params(self):
        """Multi-GPU: make the Gamma shape/rate tables (only needed as outputs) complete on every rank."""
        if self.xu is None:
            return
        if self.hyper:
            self.xu.gather(self.rate_xi, self.E_xi)
            self.xi_.gather(self.rate_eta, self.E_eta)
        if self.keep_params:
            self.xu.gather(self.shp_theta, self.rte_theta)
            self.xi_.gather(self.shp_beta, self.rte_beta)

    # -- a11 extras (single GPU; parity unpinned) ------------------------------------------------
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
        """One sweep with the multinomial (digamma) allocation of docs/Models.tex:652-664."""
        if self.r.world > 1:
            raise NotImplementedError("digamma allocation is single-GPU in this round")
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
        """Observed-only HPF ELBO and its six components (pmf_hpf_elbo); one D2H of 6 doubles."""
        if self.r.world > 1:
            raise NotImplementedError("ELBO is single-GPU in this round")
        if refresh_geomean or getattr(self, "G_theta", None) is None:
            self.geomean_tables()
        out = torch.zeros(6, dtype=torch.float64, device=self.dev)
        with torch.cuda.device(self.dev):
            _cabi.call("pmf_hpf_elbo", self.r.by_user.handle, self.K, self.ld, self.E_theta.data_ptr(), self.E_beta.data_ptr(),
                       self.G_theta.data_ptr(), self.G_beta.data_ptr(), self.shp_theta.data_ptr(), self.rte_theta.data_ptr(),
                       self.shp_beta.data_ptr(), self.rte_beta.data_ptr(), self.rate_xi.data_ptr(), self.rate_eta.data_ptr(),
                       0, self.N, 0, self.M, cfg.a, cfg.a_prime, cfg.b_prime, cfg.c, cfg.c_prime, cfg.d_prime,
                       out.data_ptr(), _cabi.stream_ptr())
        parts = out.cpu().numpy()
        return float(parts.sum()), parts

    # -- accounting ----------------------------------------------------------------------------
    def algorithmic_bytes_per_sweep(self):
        """per pass nnz*(4K+8) + R*(16K+4) (+12 R for the HPF hyper vectors)."""
        K, nnz = self.K, self.r.nnz
        per_row = 16 * K + 4 + (12 if self.hyper is not None else 0)
        return 2 * nnz * (4 * K + 8) + (self.N + self.M) * per_row
