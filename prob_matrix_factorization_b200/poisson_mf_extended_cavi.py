"""Drop-in for the reference's ``src.models.poisson_mf_extended_cavi`` (poisson_mf_extended_cavi.py:8-264).

Model: x_ui ~ Poisson(phi_u * psi_i * theta_u . beta_i) with Gamma(a0, b0) priors on everything.  Same
``PoissonMFExtendedCAVIConfig`` fields/defaults and the same ``PoissonMFExtendedCAVI`` surface (``fit / predict /
evaluate_rmse`` and the public ``a_* / b_* / E_*`` arrays for theta, beta, phi, psi); the sweeps run on a B200:

* ``_build_index_lists`` (:78-87)                    -> ``pmf_csr_build``
* user loop (:110-164) / item loop (:169-216)        -> ``pmf_gamma_pass_ext`` (the Gamma-Poisson pass kernel, MODE 2)
* ``predict`` / ``evaluate_rmse`` (:239-264)         -> ``pmf_scale_rows`` + ``pmf_predict`` / ``pmf_eval_stats``

Reference behaviour kept on purpose: the allocation divides by the raw dot product (its clamped ``rate_est`` is never
used, :137-142), the scalar's rate uses the row's freshly updated mean (:160-164), and a row without observations gets
prior shape/rate but keeps the expectations it was initialised with (:112-118).  The in-row second pass over the
gathered rows is not needed: sum_t psi_t (beta_t . theta_new) = theta_new . sum_t psi_t beta_t, and that sum is the rate
update the kernel already holds.  Single GPU (a "next" row, not part of the multi-GPU headline).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _cabi
from ._engine import EvalSet, eval_stats, normalise_ids, pad_table, predict, row_stride, table_to_host
from .poisson_mf_cavi import _DeviceBacked
from .ratings import DEFAULT_SEG_LEN, DeviceRatings, to_device


@dataclass
class PoissonMFExtendedCAVIConfig:
    n_factors: int = 20
    a0: float = 0.3
    b0: float = 1.0
    max_iter: int = 100
    tol: Optional[float] = 1e-4
    random_state: int = 42
    verbose: bool = True


class ExtendedEngine:
    """Device state of the extended model and its sweep (user pass, then item pass with the new theta / phi)."""

    def __init__(self, ratings: DeviceRatings, K, a0, b0):
        if ratings.world > 1:
            raise NotImplementedError("the extended Poisson model is single-GPU")
        self.r, self.dev, self.K, self.ld = ratings, ratings.device, int(K), row_stride(K)
        self.a0, self.b0 = float(a0), float(b0)
        self.N, self.M = ratings.n_users, ratings.n_items
        tab = lambda rows: torch.zeros((rows, self.ld), dtype=torch.float32, device=self.dev)
        vec = lambda rows: torch.zeros(rows, dtype=torch.float32, device=self.dev)
        self.E_theta, self.shp_theta, self.rte_theta = tab(self.N), tab(self.N), tab(self.N)
        self.E_beta, self.shp_beta, self.rte_beta = tab(self.M), tab(self.M), tab(self.M)
        self.E_phi, self.shp_phi, self.rte_phi = vec(self.N), vec(self.N), vec(self.N)
        self.E_psi, self.shp_psi, self.rte_psi = vec(self.M), vec(self.M), vec(self.M)
        self.ws_user = ratings.by_user.workspace(self.ld)
        self.ws_item = ratings.by_item.workspace(self.ld)
        self._scaled = None

    def load_means(self, E_theta, E_beta, E_phi, E_psi):
        self.E_theta.copy_(pad_table(E_theta, self.ld, self.dev))
        self.E_beta.copy_(pad_table(E_beta, self.ld, self.dev))
        self.E_phi.copy_(to_device(np.asarray(E_phi, dtype=np.float32), self.dev))
        self.E_psi.copy_(to_device(np.asarray(E_psi, dtype=np.float32), self.dev))
        self._scaled = None

    def _pass(self, grouped, E_oth, s_oth, E_self, shp, rte, s_shp, s_rte, s_mean, ws):
        _cabi.call("pmf_gamma_pass_ext", grouped.handle, self.K, self.ld, E_oth.data_ptr(), s_oth.data_ptr(),
                   E_self.data_ptr(), shp.data_ptr(), rte.data_ptr(), s_shp.data_ptr(), s_rte.data_ptr(),
                   s_mean.data_ptr(), self.a0, self.b0, _cabi.ptr(ws), _cabi.stream_ptr())

    def sweep(self):
        with torch.cuda.device(self.dev):
            self._pass(self.r.by_user, self.E_beta, self.E_psi, self.E_theta, self.shp_theta, self.rte_theta,
                       self.shp_phi, self.rte_phi, self.E_phi, self.ws_user)
            self._pass(self.r.by_item, self.E_theta, self.E_phi, self.E_beta, self.shp_beta, self.rte_beta,
                       self.shp_psi, self.rte_psi, self.E_psi, self.ws_item)
        self._scaled = None

    def scaled_tables(self):
        """(phi * E_theta, psi * E_beta): predictions are then a plain row dot product."""
        if self._scaled is None:
            Fu, Fi = torch.empty_like(self.E_theta), torch.empty_like(self.E_beta)
            with torch.cuda.device(self.dev):
                _cabi.call("pmf_scale_rows", self.E_theta.data_ptr(), self.E_phi.data_ptr(), self.N, self.ld, Fu.data_ptr(),
                           _cabi.stream_ptr())
                _cabi.call("pmf_scale_rows", self.E_beta.data_ptr(), self.E_psi.data_ptr(), self.M, self.ld, Fi.data_ptr(),
                           _cabi.stream_ptr())
            self._scaled = (Fu, Fi)
        return self._scaled

    def close(self):
        pass


class PoissonMFExtendedCAVI(_DeviceBacked):
    """
    Extended Poisson Matrix Factorization with mean-field VI (CAVI updates), B200 engine.
    Model: x_ij ~ Poisson(phi_u * psi_i * (theta_u^T beta_i))
    """

    _tables = {"a_theta": "shp_theta", "b_theta": "rte_theta", "E_theta": "E_theta",
               "a_beta": "shp_beta", "b_beta": "rte_beta", "E_beta": "E_beta"}
    _vectors = {"a_phi": "shp_phi", "b_phi": "rte_phi", "E_phi": "E_phi",
                "a_psi": "shp_psi", "b_psi": "rte_psi", "E_psi": "E_psi"}
    _table_names = tuple(_tables) + tuple(_vectors)

    def __init__(self, config: PoissonMFExtendedCAVIConfig, device=None, seg_len=DEFAULT_SEG_LEN):
        self._init_backing()
        self.config = config
        self.n_users = None
        self.n_items = None
        self._device = device
        self._seg_len = seg_len
        self._init = None
        self.n_iter_ = 0
        self.val_rmse_history_ = []

    def _infer_dimensions(self, train_df):
        self.n_users = int(train_df["u"].max()) + 1      # :47-51
        self.n_items = int(train_df["i"].max()) + 1
        if self.config.verbose:
            print(f"Inferred n_users={self.n_users}, n_items={self.n_items}")

    def _initial_state(self):
        """Host draws in the reference's order (:54-76): a_theta, a_beta, a_phi, a_psi; every rate = b0."""
        cfg = self.config
        rng = np.random.default_rng(cfg.random_state)
        K = cfg.n_factors
        st = {"a_theta": cfg.a0 + rng.gamma(1.0, 0.1, size=(self.n_users, K)),
              "a_beta": cfg.a0 + rng.gamma(1.0, 0.1, size=(self.n_items, K)),
              "a_phi": cfg.a0 + rng.gamma(1.0, 0.1, size=self.n_users),
              "a_psi": cfg.a0 + rng.gamma(1.0, 0.1, size=self.n_items),
              "b_theta": cfg.b0 * np.ones((self.n_users, K)), "b_beta": cfg.b0 * np.ones((self.n_items, K)),
              "b_phi": cfg.b0 * np.ones(self.n_users), "b_psi": cfg.b0 * np.ones(self.n_items)}
        for f in ("theta", "beta", "phi", "psi"):
            st["E_" + f] = st["a_" + f] / st["b_" + f]
        return st

    def _materialise(self, name):
        eng = self._engine
        if eng is None:
            return None
        if self.n_iter_ == 0 and not name.startswith("E_"):
            return self._init[name]                      # no sweep has run: shape/rate are still the initial values
        if name in self._tables:
            return table_to_host(getattr(eng, self._tables[name]), self.config.n_factors)
        return getattr(eng, self._vectors[name]).to(torch.float64).cpu().numpy()

    # -- training -----------------------------------------------------------------------------------
    def fit(self, train_df, val_df=None):
        """Same contract as poisson_mf_extended_cavi.py:89-237."""
        self._infer_dimensions(train_df)
        val = None
        if val_df is not None:
            val = (val_df["u"].to_numpy(), val_df["i"].to_numpy(), val_df["rating"].to_numpy())
        return self.fit_arrays(train_df["u"].to_numpy(), train_df["i"].to_numpy(), train_df["rating"].to_numpy(), None, val)

    def fit_arrays(self, user_ids, item_ids, ratings, init=None, val=None):
        _cabi.require_cuda()
        cfg = self.config
        if self.n_users is None:
            self.n_users, self.n_items = int(np.max(user_ids)) + 1, int(np.max(item_ids)) + 1
        self._init = self._initial_state() if init is None else init
        dr = DeviceRatings(user_ids, item_ids, ratings, self.n_users, self.n_items, self._device, seg_len=self._seg_len)
        eng = ExtendedEngine(dr, cfg.n_factors, cfg.a0, cfg.b0)
        eng.load_means(self._init["E_theta"], self._init["E_beta"], self._init["E_phi"], self._init["E_psi"])
        self._engine = eng
        self._invalidate()
        self.n_iter_ = 0
        self.val_rmse_history_ = []
        ev = EvalSet(val[0], val[1], val[2], self.n_users, self.n_items, eng.dev) if val is not None else None
        prev_val_rmse = None
        for it in range(1, cfg.max_iter + 1):
            if cfg.verbose:
                print(f"\nCAVI iteration {it}/{cfg.max_iter}")
            eng.sweep()
            self.n_iter_ = it
            if ev is not None:
                val_rmse = self._eval(ev)["rmse"]
                self.val_rmse_history_.append(val_rmse)
                if cfg.verbose:
                    print(f"Validation RMSE: {val_rmse:.4f}")
                if prev_val_rmse is not None:
                    improvement = prev_val_rmse - val_rmse
                    if cfg.verbose:
                        print(f"Improvement: {improvement:.6f}")
                    if cfg.tol is not None and improvement < cfg.tol:      # :229
                        if cfg.verbose:
                            print("Early stopping.")
                        break
                prev_val_rmse = val_rmse
        self._invalidate()
        return self

    # -- prediction / evaluation ----------------------------------------------------------------------
    def _eval(self, ev):
        e = self._engine
        Fu, Fi = e.scaled_tables()
        return eval_stats(ev, Fu, Fi, self.n_users, self.n_items, e.K, e.ld)

    def predict(self, user_ids, item_ids):
        """phi_u psi_i theta_u . beta_i; unseen ids give 0 (:239-258)."""
        e = self._engine
        if e is None:
            raise RuntimeError("fit() must be called before predict()")
        u = to_device(normalise_ids(user_ids, self.n_users), e.dev)
        i = to_device(normalise_ids(item_ids, self.n_items), e.dev)
        Fu, Fi = e.scaled_tables()
        return predict(u, i, Fu, Fi, self.n_users, self.n_items, e.K, e.ld)

    def evaluate_rmse(self, df):
        e = self._engine
        ev = EvalSet(df["u"].to_numpy(), df["i"].to_numpy(), df["rating"].to_numpy(), self.n_users, self.n_items, e.dev)
        return self._eval(ev)["rmse"]
