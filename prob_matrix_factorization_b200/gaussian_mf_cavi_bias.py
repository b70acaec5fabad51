"""Drop-in for the reference's ``src.models.gaussian_mf_cavi_bias`` (gaussian_mf_cavi_bias.py:9-347).

Gaussian matrix factorisation with user/item biases, mean-field CAVI.  Same config dataclass and
class surface; the four row loops per iteration (:132-165, :170-201, :206-232, :237-263) run on a
B200 as ``pmf_gauss_factor_pass`` (gather + accumulate of E[b b^T], float64 Cholesky inverse in shared
memory) and ``pmf_gauss_bias_pass``.  State is float32 on the device (covariances as packed lower
triangles) and materialises as the reference's float64 arrays -- ``V_theta`` as (N, K, K) -- on access.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _cabi
from ._engine import (DeviceLoop, EvalSet, device_loop_enabled, eval_stats, eval_stats_launch, normalise_ids, pad_table,
                      predict, row_stride, table_to_host)
from .poisson_mf_cavi import _DeviceBacked
from .ratings import DEFAULT_SEG_LEN, DeviceRatings, to_device


@dataclass
class GaussianMFCAVIConfig:
    n_factors: int = 10          # K (latent dimension)
    sigma2: float = 1.0          # observation noise variance σ²
    eta_theta2: float = 1.0      # prior variance for user factors η_θ²
    eta_beta2: float = 1.0       # prior variance for item factors η_β²
    eta_bias2: float = 1.0       # prior variance for biases η_b²
    max_iter: int = 20           # maximum CAVI iterations
    tol: float = 1e-3            # tolerance for validation RMSE improvement
    random_state: int = 42
    verbose: bool = True


def packed_stride(K):
    return _cabi.load().pmf_gauss_packed_stride(int(K))


def pack_lower(full, ldq):
    """(R, K, K) symmetric -> (R, ldq) float32 packed lower triangles (element (i,j), i>=j, at i(i+1)/2+j)."""
    R, K, _ = full.shape
    il, jl = np.tril_indices(K)
    out = np.zeros((R, ldq), dtype=np.float32)
    out[:, :len(il)] = full[:, il, jl]
    return out


def unpack_lower(packed, K):
    """(R, ldq) packed -> (R, K, K) float64 symmetric."""
    packed = np.asarray(packed, dtype=np.float64)
    R = packed.shape[0]
    il, jl = np.tril_indices(K)
    out = np.zeros((R, K, K))
    out[:, il, jl] = packed[:, :len(il)]
    out[:, jl, il] = packed[:, :len(il)]
    return out


class GaussEngine:
    """Device state of the Gaussian model for one shard of the ratings."""

    def __init__(self, ratings: DeviceRatings, K, bias):
        self.r, self.dev, self.K = ratings, ratings.device, int(K)
        self.ld, self.ldq = row_stride(K), packed_stride(K)
        self.N, self.M = ratings.n_users, ratings.n_items
        self.bias = bool(bias)
        f = lambda rows, w: torch.zeros((rows, w), dtype=torch.float32, device=self.dev)
        self.m_theta, self.m_beta = f(self.N, self.ld), f(self.M, self.ld)
        self.V_theta, self.V_beta = f(self.N, self.ldq), f(self.M, self.ldq)
        self.Q_theta, self.Q_beta = f(self.N, self.ldq), f(self.M, self.ldq)
        self.b_user = torch.zeros(self.N, dtype=torch.float32, device=self.dev) if bias else None
        self.b_item = torch.zeros(self.M, dtype=torch.float32, device=self.dev) if bias else None
        ws = lambda g: torch.empty(max(_cabi.load().pmf_gauss_workspace_bytes(g.handle, self.K), 16) // 4,
                                   dtype=torch.float32, device=self.dev)
        self.ws_user = ws(ratings.by_user) if ratings.by_user is not None else None
        self.ws_item = ws(ratings.by_item)
        self.world, self.rank = ratings.world, ratings.rank
        if self.world > 1:
            # ratings sharded along user ranges (ratings.py): the user passes are local; an item's ratings live on every
            # rank, so the item passes all-reduce the per-item statistics (NCCL) between accumulation and row update
            import torch.distributed as dist
            self.item_sums = f(self.M, self.ldq + self.ld)
            self.item_resid = torch.zeros(self.M, dtype=torch.float64, device=self.dev)
            counts = torch.from_numpy(np.diff(ratings.by_item.row_ptr()).astype(np.int32)).to(self.dev)
            dist.all_reduce(counts)
            self.item_counts = counts
        self.launches_per_sweep = 8 if bias else 4

    def load(self, m_theta, m_beta):
        """Initial state: given means, V = I, biases 0 (gaussian_mf_cavi_bias.py:52-67)."""
        K = self.K
        self.m_theta.copy_(pad_table(m_theta, self.ld, self.dev))
        self.m_beta.copy_(pad_table(m_beta, self.ld, self.dev))
        eye = pack_lower(np.eye(K)[None], self.ldq)
        for V, Q, m in ((self.V_theta, self.Q_theta, m_theta), (self.V_beta, self.Q_beta, m_beta)):
            V.copy_(torch.from_numpy(eye).to(self.dev).expand_as(V))
            il, jl = np.tril_indices(K)
            q = np.zeros((m.shape[0], self.ldq), dtype=np.float32)
            q[:, :len(il)] = eye[0, :len(il)] + m[:, il] * m[:, jl]
            Q.copy_(torch.from_numpy(q).to(self.dev))

    def sweep(self, sigma2, eta_theta2, eta_beta2, eta_bias2):
        c = _cabi.call
        st = _cabi.stream_ptr
        p = _cabi.ptr
        if self.world > 1:
            return self._sweep_sharded(sigma2, eta_theta2, eta_beta2, eta_bias2)
        with torch.cuda.device(self.dev):
            c("pmf_gauss_factor_pass", self.r.by_user.handle, self.K, p(self.m_beta), p(self.Q_beta), p(self.b_item),
              p(self.m_theta), p(self.V_theta), p(self.Q_theta), p(self.b_user), sigma2, eta_theta2, p(self.ws_user), st())
            c("pmf_gauss_factor_pass", self.r.by_item.handle, self.K, p(self.m_theta), p(self.Q_theta), p(self.b_user),
              p(self.m_beta), p(self.V_beta), p(self.Q_beta), p(self.b_item), sigma2, eta_beta2, p(self.ws_item), st())
            if self.bias:
                c("pmf_gauss_bias_pass", self.r.by_user.handle, self.K, p(self.m_beta), p(self.m_theta), p(self.b_item),
                  p(self.b_user), sigma2, eta_bias2, p(self.ws_user), st())
                c("pmf_gauss_bias_pass", self.r.by_item.handle, self.K, p(self.m_theta), p(self.m_beta), p(self.b_user),
                  p(self.b_item), sigma2, eta_bias2, p(self.ws_item), st())

    def _sweep_sharded(self, sigma2, eta_theta2, eta_beta2, eta_bias2):
        """The same four passes on one rank's user range: user-side passes are local, item-side passes are split around an
        NCCL all-reduce of the per-item statistics (row order theta -> beta -> b_u -> b_i as in :132-263)."""
        import torch.distributed as dist
        c, st, p = _cabi.call, _cabi.stream_ptr, _cabi.ptr
        bu, bi = self.r.by_user, self.r.by_item
        with torch.cuda.device(self.dev):
            if bu is not None:
                c("pmf_gauss_factor_pass", bu.handle, self.K, p(self.m_beta), p(self.Q_beta), p(self.b_item),
                  p(self.m_theta), p(self.V_theta), p(self.Q_theta), p(self.b_user), sigma2, eta_theta2, p(self.ws_user), st())
            args = (bi.handle, self.K, p(self.m_theta), p(self.Q_theta), p(self.b_user), p(self.m_beta), p(self.V_beta),
                    p(self.Q_beta), p(self.b_item), sigma2, eta_beta2, p(self.ws_item), p(self.item_sums), p(self.item_counts))
            c("pmf_gauss_factor_pass_sharded", *args, 1, st())
            dist.all_reduce(self.item_sums)
            c("pmf_gauss_factor_pass_sharded", *args, 2, st())
            if self.bias:
                if bu is not None:
                    c("pmf_gauss_bias_pass", bu.handle, self.K, p(self.m_beta), p(self.m_theta), p(self.b_item),
                      p(self.b_user), sigma2, eta_bias2, p(self.ws_user), st())
                args = (bi.handle, self.K, p(self.m_theta), p(self.m_beta), p(self.b_user), p(self.b_item), sigma2, eta_bias2,
                        p(self.ws_item), p(self.item_resid), p(self.item_counts))
                c("pmf_gauss_bias_pass_sharded", *args, 1, st())
                dist.all_reduce(self.item_resid)
                c("pmf_gauss_bias_pass_sharded", *args, 2, st())

    def eval_range(self):
        return None if self.world == 1 else (self.r.user_lo, self.r.user_hi, self.rank == self.world - 1)

    def sync(self):
        """Several GPUs, end of fit: the user-side tables live with their owners while the sweeps run -- gather them."""
        if self.world == 1:
            return
        from .parallel import RowExchange
        xu = RowExchange(self.r.user_bounds)
        xu.gather(self.m_theta, self.V_theta, self.Q_theta)
        if self.bias:
            xu.gather(self.b_user)


class GaussianMFCAVI(_DeviceBacked):
    """
    Gaussian Matrix Factorization with mean-field VI (CAVI updates), B200 engine.
    Model: r_ij ~ N(mu + b_i + b_j + theta_i^T beta_j, sigma^2)
    """

    _table_names = ("m_theta", "V_theta", "m_beta", "V_beta", "m_user_bias", "m_item_bias")
    _with_bias = True

    def __init__(self, config, device=None, seg_len=DEFAULT_SEG_LEN, shard=None):
        """``shard=(rank, world)``: one process per GPU, this rank keeps the ratings of one user range (ratings.py); the
        per-item statistics of the item passes are all-reduced over NCCL, every rank ends with the complete state."""
        self._init_backing()
        self._shard = shard
        self.config = config
        self.n_users = None
        self.n_items = None
        self.global_mean = 0.0
        self._device = device
        self._seg_len = seg_len
        self.n_iter_ = 0
        self.val_rmse_history_ = []

    def _infer_dimensions(self, train_df):
        self.n_users = int(train_df["u"].max()) + 1
        self.n_items = int(train_df["i"].max()) + 1
        if self.config.verbose:
            print(f"Inferred n_users={self.n_users}, n_items={self.n_items}")

    def _initial_state(self):
        """Host draws in the reference's order (gaussian_mf_cavi_bias.py:52-58)."""
        rng = np.random.default_rng(self.config.random_state)
        K = self.config.n_factors
        return {"m_theta": 0.1 * rng.standard_normal((self.n_users, K)),
                "m_beta": 0.1 * rng.standard_normal((self.n_items, K))}

    def _materialise(self, name):
        e = self._engine
        if e is None:
            return None
        K = self.config.n_factors
        if name in ("m_theta", "m_beta"):
            return table_to_host(getattr(e, name), K)
        if name in ("V_theta", "V_beta"):
            return unpack_lower(getattr(e, name).cpu().numpy(), K)
        if not self._with_bias:
            raise AttributeError(name)
        t = e.b_user if name == "m_user_bias" else e.b_item
        return t.double().cpu().numpy()

    def fit(self, train_df, val_df=None, global_mean=0.0):
        """Run CAVI on (pre-centred) training data; same contract as gaussian_mf_cavi_bias.py:91-286."""
        self.global_mean = global_mean
        self._infer_dimensions(train_df)
        init = self._initial_state()
        val = None
        if val_df is not None:
            val = (val_df["u"].to_numpy(), val_df["i"].to_numpy(), val_df["rating"].to_numpy())
        return self.fit_arrays(train_df["u"].to_numpy(), train_df["i"].to_numpy(), train_df["rating"].to_numpy(),
                               init, val, global_mean)

    def fit_arrays(self, user_ids, item_ids, ratings, init=None, val=None, global_mean=0.0):
        _cabi.require_cuda()
        cfg = self.config
        self.global_mean = global_mean
        if self.n_users is None:
            self.n_users, self.n_items = int(np.max(user_ids)) + 1, int(np.max(item_ids)) + 1
        if init is None:
            init = self._initial_state()
        dr = DeviceRatings(user_ids, item_ids, ratings, self.n_users, self.n_items, self._device, seg_len=self._seg_len,
                           shard=self._shard, item_chunks=1)
        eng = GaussEngine(dr, cfg.n_factors, self._with_bias)
        eng.load(init["m_theta"], init["m_beta"])
        self._engine = eng
        self._invalidate()
        self.n_iter_ = 0
        self.val_rmse_history_ = []
        ev = None
        if val is not None:
            ev = EvalSet(val[0], val[1], val[2], self.n_users, self.n_items, eng.dev, drop_invalid=True,
                         user_range=eng.eval_range())
        eta_bias2 = getattr(cfg, "eta_bias2", 1.0)
        prev_val_rmse = None
        host_iters = range(1, cfg.max_iter + 1)
        if device_loop_enabled() and not cfg.verbose and cfg.max_iter >= 1 and eng.world == 1:
            host_iters = self._fit_on_device(eng, ev, eta_bias2)
        for it in host_iters:
            if cfg.verbose:
                print(f"\nCAVI iteration {it}/{cfg.max_iter}")
            eng.sweep(cfg.sigma2, cfg.eta_theta2, cfg.eta_beta2, eta_bias2)
            self.n_iter_ = it
            if ev is not None:
                st = self._eval(ev)
                if st["count"] == 0:
                    print("Warning: No valid (u,i) pairs.")
                val_rmse = st["rmse"]
                self.val_rmse_history_.append(val_rmse)
                if cfg.verbose:
                    if self._with_bias:
                        print(f"Validation RMSE: {val_rmse:.4f} | MacroMAE: {st['macro_mae']:.4f}")
                    else:
                        print(f"Validation RMSE: {val_rmse:.4f}")
                if prev_val_rmse is not None:
                    improvement = prev_val_rmse - val_rmse
                    if cfg.verbose:
                        print(f"Improvement: {improvement:.6f}")
                    if improvement >= 0 and improvement < cfg.tol:          # :279 (needs 0 <= imp < tol)
                        if cfg.verbose:
                            print("Early stopping: small improvement on validation.")
                        break
                prev_val_rmse = val_rmse
        eng.sync()
        self._invalidate()
        return self

    def _fit_on_device(self, eng, ev, eta_bias2):
        """The whole loop (4 passes + validation statistics + the stopping rule 0 <= improvement < tol, :279) as one
        CUDA graph with a device-side WHILE (DeviceLoop).  Returns the iterations left for the host loop."""
        cfg = self.config
        try:
            loop = DeviceLoop(eng.dev, cfg.max_iter, ev_out=None if ev is None else ev.out, rule=1,
                              tol=None if ev is None else cfg.tol)
            with loop.body():
                eng.sweep(cfg.sigma2, cfg.eta_theta2, cfg.eta_beta2, eta_bias2)
                if ev is not None:
                    eval_stats_launch(ev, eng.m_theta, eng.m_beta, self.n_users, self.n_items, eng.K, eng.ld, eng.b_user,
                                      eng.b_item, self.global_mean)
        except _cabi.PMFError as exc:
            if exc.status != _cabi.PMF_EUNSUPPORTED:
                raise
            return range(1, cfg.max_iter + 1)
        self.n_iter_, hist = loop.run()
        loop.free()
        self.val_rmse_history_ = [float(v) for v in hist]
        for v in self.val_rmse_history_:
            if np.isnan(v):
                print("Warning: No valid (u,i) pairs.")
        return range(0)

    def _eval(self, ev):
        e = self._engine
        return eval_stats(ev, e.m_theta, e.m_beta, self.n_users, self.n_items, e.K, e.ld, e.b_user, e.b_item,
                          self.global_mean)

    def predict(self, user_ids, item_ids, global_mean=0.0):
        """b_i + b_j + m_θi^T m_βj + global_mean; unseen ids give global_mean (:291-316)."""
        e = self._engine
        if e is None:
            raise RuntimeError("fit() must be called before predict()")
        u = to_device(normalise_ids(user_ids, self.n_users), e.dev)
        i = to_device(normalise_ids(item_ids, self.n_items), e.dev)
        # the kernel adds the mean in float32; add it here in float64 like the reference (:316)
        return predict(u, i, e.m_theta, e.m_beta, self.n_users, self.n_items, e.K, e.ld, e.b_user, e.b_item, 0.0) + global_mean

    def _frame_eval(self, df, global_mean):
        e = self._engine
        ev = EvalSet(df["u"].to_numpy(), df["i"].to_numpy(), df["rating"].to_numpy(dtype=float), self.n_users,
                     self.n_items, e.dev, drop_invalid=True)
        return eval_stats(ev, e.m_theta, e.m_beta, self.n_users, self.n_items, e.K, e.ld, e.b_user, e.b_item, global_mean)

    def evaluate_rmse(self, df, global_mean):
        """RMSE on the original scale, ignoring unseen users/items (:318-333)."""
        st = self._frame_eval(df, global_mean)
        if st["count"] == 0:
            print("Warning: No valid (u,i) pairs.")
            return np.nan
        return st["rmse"]

    def evaluate_macro_mae(self, df, global_mean):
        st = self._frame_eval(df, global_mean)
        return np.nan if st["count"] == 0 else st["macro_mae"]

    def log_predictive_likelihood(self, df, sigma=None):
        """``GaussianLogPredictiveLikelihood(df, m_theta, m_beta, sigma)`` (metrics.py:18-35) evaluated on the device.

        As the reference's callers use it (run_gaussian_mf_best_k.py:54): predictions are the factor means' dot product
        only (no biases, no global mean), ``sigma`` defaults to ``config.sigma2`` and is SQUARED by the function
        although it already is a variance (metrics.py:33).  Rows with unseen ids (an IndexError in the reference) are
        dropped."""
        e = self._engine
        if e is None:
            raise RuntimeError("fit() must be called before log_predictive_likelihood()")
        sigma = self.config.sigma2 if sigma is None else sigma
        ev = EvalSet(df["u"].to_numpy(), df["i"].to_numpy(), df["rating"].to_numpy(dtype=float), self.n_users,
                     self.n_items, e.dev, drop_invalid=True)
        st = eval_stats(ev, e.m_theta, e.m_beta, self.n_users, self.n_items, e.K, e.ld)
        variance = float(sigma) ** 2
        return float(-0.5 * st["count"] * np.log(2 * np.pi * variance) - st["sse"] / (2 * variance))
