"""Drop-in for the reference's ``src.models.hpf_cavi`` (hpf_cavi.py:7-241).

Observed-only hierarchical Poisson factorisation with CAVI updates.  Same config dataclass and
class surface; the per-row loops (hpf_cavi.py:126-151, :162-185) and the xi / eta rate updates
(:158, :192) run fused in ``pmf_gamma_pass`` on a B200.  The allocation uses arithmetic means,
exactly as the reference code does (NOT the digamma form of docs/Models.tex).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _cabi
from ._engine import (DeviceLoop, EvalSet, GammaEngine, Trace, device_loop_enabled, eval_stats, eval_stats_launch,
                      normalise_ids, predict, row_stride, table_to_host)
from .host_draws import divide, gamma_shape1
from .poisson_mf_cavi import _DeviceBacked
from .ratings import DEFAULT_SEG_LEN, DeviceRatings, to_device


@dataclass
class HPF_CAVI_Config:
    n_factors: int = 20
    a: float = 0.3              # Shape for theta
    a_prime: float = 0.3        # Shape for xi (user rate prior)
    b_prime: float = 1.0        # Rate for xi
    c: float = 0.3              # Shape for beta
    c_prime: float = 0.3        # Shape for eta (item rate prior)
    d_prime: float = 1.0        # Rate for eta
    max_iter: int = 100
    tol: Optional[float] = 1e-4
    random_state: int = 42
    verbose: bool = True


class HPF_CAVI(_DeviceBacked):
    """
    Hierarchical Poisson Factorization with CAVI updates on OBSERVED data only (B200 engine).

        x_ui ~ Poisson(theta_u^T beta_i);  theta_uk ~ Gamma(a, xi_u);  xi_u ~ Gamma(a', b')
        beta_ik ~ Gamma(c, eta_i);  eta_i ~ Gamma(c', d')
    """

    _table_names = ("gamma_a_theta", "gamma_b_theta", "gamma_a_beta", "gamma_b_beta", "gamma_b_xi", "gamma_b_eta",
                    "E_theta", "E_beta", "E_xi", "E_eta")
    _dev_names = {"gamma_a_theta": "shp_theta", "gamma_b_theta": "rte_theta", "gamma_a_beta": "shp_beta",
                  "gamma_b_beta": "rte_beta", "gamma_b_xi": "rate_xi", "gamma_b_eta": "rate_eta",
                  "E_theta": "E_theta", "E_beta": "E_beta", "E_xi": "E_xi", "E_eta": "E_eta"}

    def __init__(self, config: HPF_CAVI_Config, device=None, shard=None, seg_len=DEFAULT_SEG_LEN,
                 allocation="mean", track_elbo=False):
        """``allocation="mean"`` is the reference code's update (arithmetic means, hpf_cavi.py:140-151);
        ``"digamma"`` is the textbook multinomial step of docs/Models.tex:652-664 (no reference code; parity
        unpinned).  ``track_elbo`` records the ELBO after every sweep in ``elbo_history_``."""
        if allocation not in ("mean", "digamma"):
            raise ValueError("allocation must be 'mean' or 'digamma'")
        self._init_backing()
        self._allocation = allocation
        self._track_elbo = bool(track_elbo)
        self.elbo_history_ = []
        self.config = config
        self.n_users = None
        self.n_items = None
        self.gamma_a_xi = None   # scalars (hpf_cavi.py:81, :85)
        self.gamma_a_eta = None
        self._device = device
        self._shard = shard
        self._seg_len = seg_len
        self._auto_close = True
        self._ratings_kw = {}     # extra DeviceRatings arguments (tile overrides; experiments / tests)
        self._engine_kw = {}      # extra GammaEngine arguments (exchange, item_chunks)
        self.n_iter_ = 0
        self.val_rmse_history_ = []
        self._init = None

    def _infer_dimensions(self, train_df):
        self.n_users = int(train_df["u"].max()) + 1
        self.n_items = int(train_df["i"].max()) + 1
        if self.config.verbose:
            print(f"Inferred n_users={self.n_users}, n_items={self.n_items}")

    def _initial_state(self):
        """Host draws in the reference's order: a_theta, b_theta, a_beta, b_beta (hpf_cavi.py:66-89)."""
        cfg = self.config
        rng = np.random.default_rng(cfg.random_state)
        K, N, M = cfg.n_factors, self.n_users, self.n_items
        # cfg.a + rng.gamma(1.0, 0.1, size=(N, K)) etc., bit for bit (host_draws: NumPy's stream replayed by all host cores)
        a_t = gamma_shape1(rng, 0.1, (N, K), cfg.a)
        b_t = gamma_shape1(rng, 0.1, (N, K), cfg.b_prime)
        a_b = gamma_shape1(rng, 0.1, (M, K), cfg.c)
        b_b = gamma_shape1(rng, 0.1, (M, K), cfg.d_prime)
        a_xi = cfg.a_prime + K * cfg.a
        a_eta = cfg.c_prime + K * cfg.c
        b_xi = cfg.b_prime * np.ones(N)
        b_eta = cfg.d_prime * np.ones(M)
        return {"gamma_a_theta": a_t, "gamma_b_theta": b_t, "gamma_a_beta": a_b, "gamma_b_beta": b_b,
                "gamma_a_xi": a_xi, "gamma_a_eta": a_eta, "gamma_b_xi": b_xi, "gamma_b_eta": b_eta,
                "E_theta": divide(a_t, b_t), "E_beta": divide(a_b, b_b), "E_xi": a_xi / b_xi, "E_eta": a_eta / b_eta}

    def _materialise(self, name):
        eng = self._engine
        if eng is None:
            return None
        if self.n_iter_ == 0 and self._init is not None and name in self._init:
            return self._init[name]
        t = getattr(eng, self._dev_names[name])
        if t is None:
            return None
        if t.dim() == 1:
            return t.double().cpu().numpy()
        return table_to_host(t, self.config.n_factors)

    def fit(self, train_df, val_df=None):
        self._infer_dimensions(train_df)
        init = self._initial_state()
        val = None
        if val_df is not None:
            val = (val_df["u"].to_numpy(), val_df["i"].to_numpy(), val_df["rating"].to_numpy())
        return self.fit_arrays(train_df["u"].to_numpy(), train_df["i"].to_numpy(), train_df["rating"].to_numpy(),
                               init, val)

    def fit_arrays(self, user_ids, item_ids, ratings, init=None, val=None):
        """``fit`` on host arrays: H2D, device grouping, ``max_iter`` sweeps (the timed e2e path)."""
        _cabi.require_cuda()
        cfg = self.config
        if self.n_users is None:
            self.n_users, self.n_items = int(np.max(user_ids)) + 1, int(np.max(item_ids)) + 1
        if init is None:
            init = self._initial_state()
        self.gamma_a_xi, self.gamma_a_eta = init["gamma_a_xi"], init["gamma_a_eta"]
        tr = Trace()
        if self._engine is not None:
            self._engine.close()          # collective on multi-GPU runs: every rank re-fits together
        tr.mark("close previous engine")
        dr = DeviceRatings(user_ids, item_ids, ratings, self.n_users, self.n_items, self._device,
                           seg_len=self._seg_len, shard=self._shard,
                           row_bytes=None if self._allocation == "digamma" else 4 * row_stride(cfg.n_factors),
                           **self._ratings_kw)
        tr.mark("ratings H2D + routing + grouping")
        hyper = {"user_shape": float(init["gamma_a_xi"]), "user_rate_prior": float(cfg.b_prime),
                 "item_shape": float(init["gamma_a_eta"]), "item_rate_prior": float(cfg.d_prime)}
        eng = GammaEngine(dr, cfg.n_factors, cfg.a, cfg.c, None, None, hyper=hyper, **self._engine_kw)
        tr.mark("engine tables (+ symmetric memory)")
        eng.load_means(init["E_theta"], init["E_beta"], init["E_xi"], init["E_eta"])
        tr.mark("initial factors H2D")
        if self._allocation == "digamma" or self._track_elbo:
            eng.load_params(init["gamma_a_theta"], init["gamma_b_theta"], init["gamma_a_beta"], init["gamma_b_beta"],
                            init["gamma_b_xi"], init["gamma_b_eta"])
            eng.geomean_tables()
        self.elbo_history_ = []
        self._engine = eng
        self._init = init
        self._invalidate()
        self.n_iter_ = 0
        self.val_rmse_history_ = []
        ev = None
        if val is not None:
            ev = EvalSet(val[0], val[1], val[2], self.n_users, self.n_items, eng.dev, user_range=eng.eval_range())
        prev_val_rmse = None
        # the Gamma shape/rate tables are outputs only: they are written by the sweep that can be the last one
        params_every_sweep = self._track_elbo or (ev is not None and cfg.tol is not None)
        host_iters = range(1, cfg.max_iter + 1)
        if (device_loop_enabled() and not cfg.verbose and eng.world == 1 and self._allocation == "mean"
                and not self._track_elbo and cfg.max_iter >= 1):
            host_iters = self._fit_on_device(eng, ev, params_every_sweep)
        if eng.world > 1 and ev is None and not cfg.verbose and cfg.max_iter >= 1 and self._allocation == "mean" and not self._track_elbo:
            eng.sweeps(cfg.max_iter)          # software-pipelined across sweeps (combine of sweep s under user pass s+1)
            self.n_iter_ = cfg.max_iter
            host_iters = range(0)
        for it in host_iters:
            if cfg.verbose:
                print(f"\nHPF_CAVI iteration {it}/{cfg.max_iter}")
            if self._allocation == "digamma":
                eng.sweep_digamma()
            else:
                eng.sweep(write_params=params_every_sweep or it == cfg.max_iter)
            self.n_iter_ = it
            if self._track_elbo:
                self.elbo_history_.append(eng.elbo(cfg, refresh_geomean=self._allocation != "digamma")[0])
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
                    if cfg.tol is not None and improvement < cfg.tol:      # hpf_cavi.py:207
                        if cfg.verbose:
                            print("Early stopping.")
                        break
                prev_val_rmse = val_rmse
        tr.mark(f"{self.n_iter_} sweeps")
        eng.sync_params()
        tr.mark("gather shape/rate tables")
        if self._auto_close:
            eng.close()                   # symmetric-memory tables (multi-GPU) become ordinary device tensors
        if self.n_iter_ > 0:
            self._init = None
        self._invalidate()
        return self

    def _fit_on_device(self, eng, ev, params_every_sweep):
        """The whole loop as one CUDA graph (DeviceLoop): no host round trip per iteration.  Returns the iterations that
        are left for the host loop (the last sweep of a fit without early stopping, which alone writes the Gamma
        parameters; everything if the driver cannot build the graph)."""
        cfg = self.config
        in_graph = cfg.max_iter if params_every_sweep else cfg.max_iter - 1
        if in_graph >= 1:
            tr = Trace()
            try:
                loop = DeviceLoop(eng.dev, in_graph, ev_out=None if ev is None else ev.out, rule=0, tol=cfg.tol)
                tr.mark("device loop: construct")
                with loop.body():
                    eng.sweep(write_params=params_every_sweep)
                    if ev is not None:
                        eval_stats_launch(ev, eng.E_theta, eng.E_beta, self.n_users, self.n_items, eng.K, eng.ld)
            except _cabi.PMFError as exc:
                if exc.status != _cabi.PMF_EUNSUPPORTED:
                    raise
                return range(1, cfg.max_iter + 1)
            tr.mark("device loop: capture + instantiate")
            self.n_iter_, hist = loop.run()
            tr.mark(f"device loop: run ({self.n_iter_} iterations)")
            self.val_rmse_history_ = [float(v) for v in hist]
            loop.free()
            tr.mark("device loop: free")
        return range(in_graph + 1, cfg.max_iter + 1)

    def elbo(self, return_parts=False):
        """Evidence lower bound of the current variational state (observed-only HPF; parity unpinned)."""
        e = self._engine
        if e is None:
            raise RuntimeError("fit() must be called before elbo()")
        if self.n_iter_ == 0 and self._init is not None and not (self._allocation == "digamma" or self._track_elbo):
            i = self._init
            e.load_params(i["gamma_a_theta"], i["gamma_b_theta"], i["gamma_a_beta"], i["gamma_b_beta"],
                          i["gamma_b_xi"], i["gamma_b_eta"])
        total, parts = e.elbo(self.config, refresh_geomean=True)
        return (total, parts) if return_parts else total

    def _eval(self, ev):
        e = self._engine
        return eval_stats(ev, e.E_theta, e.E_beta, self.n_users, self.n_items, e.K, e.ld)

    def predict(self, user_ids, item_ids):
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
