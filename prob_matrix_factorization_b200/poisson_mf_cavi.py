"""Drop-in for the reference's ``src.models.poisson_mf_cavi`` (poisson_mf_cavi.py:9-251).

Same ``PoissonMFCAVIConfig`` fields/defaults and the same ``PoissonMFCAVI`` surface
(``fit / predict / evaluate_rmse / evaluate_macro_mae`` and the public ``a_* / b_* / E_*``
arrays), but the CAVI sweeps run on a B200 through libpmf_b200:

* ``_build_index_lists`` (:73-84)      -> ``pmf_csr_build`` (stable device grouping, both sides)
* user / item row loops (:135-197)     -> ``pmf_gamma_pass`` (fused SDDMM + allocation + row sums + a/b)
* ``predict`` / ``evaluate_*`` (:221-251) -> ``pmf_predict`` / ``pmf_eval_stats``

Initial values are drawn on the host with NumPy's PCG64 in the reference's order (:50-71) so
runs are reproducible against it; state lives in float32 on the device and is exposed as
float64 NumPy arrays on attribute access.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _cabi
from ._engine import (DeviceLoop, EvalSet, GammaEngine, device_loop_enabled, eval_stats, eval_stats_launch, normalise_ids,
                      predict, row_stride, table_to_host)
from .host_draws import divide, gamma_shape1
from .ratings import DEFAULT_SEG_LEN, DeviceRatings, to_device


@dataclass
class PoissonMFCAVIConfig:
    n_factors: int = 20          # K (latent dimension)
    a0: float = 0.3              # Hyperparameter a for Gamma prior
    b0: float = 1.0              # Hyperparameter b for Gamma prior
    max_iter: int = 100          # Maximum CAVI iterations
    tol: Optional[float] = 1e-4  # Tolerance for convergence (None to disable)
    random_state: int = 42
    verbose: bool = True


class _DeviceBacked:
    """Public arrays that live on the GPU and materialise as float64 NumPy on first access."""

    _table_names: tuple = ()

    def _init_backing(self):
        object.__setattr__(self, "_host", {})
        object.__setattr__(self, "_engine", None)

    def __getattr__(self, name):
        # only reached when normal lookup fails
        if name.startswith("_") or name not in type(self)._table_names:
            raise AttributeError(name)
        host = self.__dict__.get("_host", {})
        if name not in host:
            host[name] = self._materialise(name)
        return host[name]

    def __setattr__(self, name, value):
        if name in type(self)._table_names:
            self.__dict__.setdefault("_host", {})[name] = value
            self._on_host_override(name)
        else:
            object.__setattr__(self, name, value)

    def _invalidate(self, *names):
        for n in names or type(self)._table_names:
            self._host.pop(n, None)

    def _materialise(self, name):
        return None

    def _on_host_override(self, name):
        pass


class PoissonMFCAVI(_DeviceBacked):
    """
    Poisson Matrix Factorization with mean-field VI (CAVI updates), B200 engine.
    Model: x_ij ~ Poisson(theta_i^T beta_j); theta_i, beta_j ~ Gamma(a0, b0).
    """

    _table_names = ("a_theta", "b_theta", "a_beta", "b_beta", "E_theta", "E_beta")
    _dev_names = {"a_theta": "shp_theta", "b_theta": "rte_theta", "a_beta": "shp_beta", "b_beta": "rte_beta",
                  "E_theta": "E_theta", "E_beta": "E_beta"}

    def __init__(self, config: PoissonMFCAVIConfig, device=None, shard=None, seg_len=DEFAULT_SEG_LEN):
        self._init_backing()
        self.config = config
        self.n_users = None
        self.n_items = None
        self._device = device
        self._shard = shard
        self._seg_len = seg_len
        self._auto_close = True
        self._ratings_kw = {}     # extra DeviceRatings arguments (tile overrides; experiments / tests)
        self._engine_kw = {}      # extra GammaEngine arguments (exchange, item_chunks)
        self.n_iter_ = 0
        self.val_rmse_history_ = []

    # -- reference helpers kept for API parity ----------------------------------------------------
    def _infer_dimensions(self, train_df):
        self.n_users = int(train_df["u"].max()) + 1      # poisson_mf_cavi.py:44-46
        self.n_items = int(train_df["i"].max()) + 1
        if self.config.verbose:
            print(f"Inferred n_users={self.n_users}, n_items={self.n_items}")

    def _initial_state(self):
        """Host draws in the reference's order (poisson_mf_cavi.py:50-71)."""
        rng = np.random.default_rng(self.config.random_state)
        K = self.config.n_factors
        # a0 + rng.gamma(1.0, 0.1, size=...), bit for bit (host_draws: NumPy's stream replayed by all host cores)
        a_theta = gamma_shape1(rng, 0.1, (self.n_users, K), self.config.a0)
        a_beta = gamma_shape1(rng, 0.1, (self.n_items, K), self.config.a0)
        return {"a_theta": a_theta, "a_beta": a_beta,
                "E_theta": divide(a_theta, float(self.config.b0)), "E_beta": divide(a_beta, float(self.config.b0))}

    def _materialise(self, name):
        eng = self._engine
        if eng is None:
            return None
        t = getattr(eng, self._dev_names[name])
        if t is None:
            return None
        if self.n_iter_ == 0 and name in ("b_theta", "b_beta"):
            rows = self.n_users if name == "b_theta" else self.n_items
            return self.config.b0 * np.ones((rows, self.config.n_factors))
        return table_to_host(t, self.config.n_factors)

    def _on_host_override(self, name):
        eng = self.__dict__.get("_engine")
        val = self._host.get(name)
        if eng is not None and val is not None and name in ("E_theta", "E_beta"):
            from ._engine import pad_table
            getattr(eng, name).copy_(pad_table(val, eng.ld, eng.dev))

    # -- training -----------------------------------------------------------------------------------
    def fit(self, train_df, val_df=None):
        """Run CAVI on the training data (same contract as poisson_mf_cavi.py:86-219)."""
        self._infer_dimensions(train_df)
        init = self._initial_state()
        user_ids = train_df["u"].to_numpy()
        item_ids = train_df["i"].to_numpy()
        ratings = train_df["rating"].to_numpy()
        val = None
        if val_df is not None:
            val = (val_df["u"].to_numpy(), val_df["i"].to_numpy(), val_df["rating"].to_numpy())
        return self.fit_arrays(user_ids, item_ids, ratings, init, val)

    def fit_arrays(self, user_ids, item_ids, ratings, init=None, val=None):
        """``fit`` on host arrays: H2D, device grouping, ``max_iter`` sweeps (the timed e2e path)."""
        _cabi.require_cuda()
        cfg = self.config
        if self.n_users is None:
            self.n_users, self.n_items = int(np.max(user_ids)) + 1, int(np.max(item_ids)) + 1
        if init is None:
            init = self._initial_state()
        if self._engine is not None:
            self._engine.close()          # collective on multi-GPU runs: every rank re-fits together
        dr = DeviceRatings(user_ids, item_ids, ratings, self.n_users, self.n_items, self._device,
                           seg_len=self._seg_len, shard=self._shard, row_bytes=4 * row_stride(cfg.n_factors),
                           **self._ratings_kw)
        eng = GammaEngine(dr, cfg.n_factors, cfg.a0, cfg.a0, cfg.b0, cfg.b0, **self._engine_kw)
        eng.load_means(init["E_theta"], init["E_beta"])
        self._engine = eng
        self._invalidate()
        self._host.update({"a_theta": init["a_theta"], "a_beta": init["a_beta"]})
        self.n_iter_ = 0
        self.val_rmse_history_ = []
        ev = None
        if val is not None:
            ev = EvalSet(val[0], val[1], val[2], self.n_users, self.n_items, eng.dev, user_range=eng.eval_range())
        prev_val_rmse = None
        # the Gamma shape/rate tables are outputs only: they are written by the sweep that can be the last one
        params_every_sweep = ev is not None and cfg.tol is not None
        host_iters = range(1, cfg.max_iter + 1)
        if device_loop_enabled() and not cfg.verbose and eng.world == 1 and cfg.max_iter >= 1:
            host_iters = self._fit_on_device(eng, ev, params_every_sweep)
        if eng.world > 1 and ev is None and not cfg.verbose and cfg.max_iter >= 1:
            eng.sweeps(cfg.max_iter)          # software-pipelined across sweeps (combine of sweep s under user pass s+1)
            self.n_iter_ = cfg.max_iter
            host_iters = range(0)
        for it in host_iters:
            if cfg.verbose:
                print(f"\nCAVI iteration {it}/{cfg.max_iter}")
            eng.sweep(write_params=params_every_sweep or it == cfg.max_iter)
            self.n_iter_ = it
            if ev is not None:
                st = self._eval(ev)
                val_rmse, val_macro_mae = st["rmse"], st["macro_mae"]
                self.val_rmse_history_.append(val_rmse)
                if cfg.verbose:
                    print(f"Validation RMSE: {val_rmse:.4f} | MacroMAE: {val_macro_mae:.4f}")
                if prev_val_rmse is not None:
                    improvement = prev_val_rmse - val_rmse
                    if cfg.verbose:
                        print(f"Improvement: {improvement:.6f}")
                    if cfg.tol is not None and improvement < cfg.tol:      # :213 (fires on negative too)
                        if cfg.verbose:
                            print("Early stopping.")
                        break
                prev_val_rmse = val_rmse
        eng.sync_params()
        if self._auto_close:
            eng.close()                   # symmetric-memory tables (multi-GPU) become ordinary device tensors
        self._invalidate()
        if self.n_iter_ == 0:
            self._host.update({"a_theta": init["a_theta"], "a_beta": init["a_beta"]})
        return self

    def _fit_on_device(self, eng, ev, params_every_sweep):
        """The whole loop as one CUDA graph (DeviceLoop): no host round trip per iteration.  Returns the iterations that
        are left for the host loop (the last sweep of a fit without early stopping, which alone writes the Gamma
        parameters; everything if the driver cannot build the graph)."""
        cfg = self.config
        in_graph = cfg.max_iter if params_every_sweep else cfg.max_iter - 1
        if in_graph >= 1:
            try:
                loop = DeviceLoop(eng.dev, in_graph, ev_out=None if ev is None else ev.out, rule=0, tol=cfg.tol)
                with loop.body():
                    eng.sweep(write_params=params_every_sweep)
                    if ev is not None:
                        eval_stats_launch(ev, eng.E_theta, eng.E_beta, self.n_users, self.n_items, eng.K, eng.ld)
            except _cabi.PMFError as exc:
                if exc.status != _cabi.PMF_EUNSUPPORTED:
                    raise
                return range(1, cfg.max_iter + 1)
            self.n_iter_, hist = loop.run()
            self.val_rmse_history_ = [float(v) for v in hist]
            loop.free()
        return range(in_graph + 1, cfg.max_iter + 1)

    # -- prediction / evaluation ----------------------------------------------------------------------
    def _eval(self, ev):
        e = self._engine
        return eval_stats(ev, e.E_theta, e.E_beta, self.n_users, self.n_items, e.K, e.ld)

    def predict(self, user_ids, item_ids):
        """Predict E[x_ij] = E[theta_i]^T E[beta_j]; unseen ids give 0 (poisson_mf_cavi.py:221-241)."""
        e = self._engine
        if e is None:
            raise RuntimeError("fit() must be called before predict()")
        u = to_device(normalise_ids(user_ids, self.n_users), e.dev)
        i = to_device(normalise_ids(item_ids, self.n_items), e.dev)
        return predict(u, i, e.E_theta, e.E_beta, self.n_users, self.n_items, e.K, e.ld)

    def _frame_eval(self, df):
        e = self._engine
        ev = EvalSet(df["u"].to_numpy(), df["i"].to_numpy(), df["rating"].to_numpy(), self.n_users, self.n_items, e.dev)
        return self._eval(ev)

    def evaluate_rmse(self, df):
        return self._frame_eval(df)["rmse"]

    def evaluate_macro_mae(self, df):
        return self._frame_eval(df)["macro_mae"]

    def log_predictive_likelihood(self, df):
        """PoissonLogPredictiveLikelihood (metrics.py:53-66) evaluated on the device."""
        return self._frame_eval(df)["poisson_lpl"]
