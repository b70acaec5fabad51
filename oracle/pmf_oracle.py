"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the per-iteration training hot path.

This file restates, in float64 NumPy, the algorithms of the reference
rogeliolopezcamara/prob-matrix-factorization model classes so that the CUDA
path can be checked against them on a GPU box where /root/reference does not
exist.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline legs may import it; the product package never does.

Pin status
----------
* CSR grouping, Poisson MF, HPF-CAVI, Gaussian MF (with and without biases),
  predict, RMSE, macro-MAE, Poisson/Gaussian LPL, HPF-MAP loss/gradients/Adam:
  **pinned** -- ``tests/golden/make_golden.py`` ran the reference's own classes
  (imported read-only from /root/reference) on seeded inputs and committed the
  outputs as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every
  function here against those files.
* ``hpf_digamma_sweeps``, ``hpf_elbo``, ``topn``: **parity unpinned** -- the
  reference has no code for the textbook digamma allocation, the ELBO or top-N
  scoring (grep finds none in src/); these follow docs/Models.tex:631-726 and
  Gopalan, Hofman & Blei (2015) and are only self-consistency checks.

Every function cites the reference file:line it follows (paths relative to
/root/reference).
"""
from __future__ import annotations

import numpy as np

RATE_FLOOR = 1e-10  # poisson_mf_cavi.py:153, hpf_cavi.py:141


# ----------------------------------------------------------------------------
# a1: observation grouping (the reference's "CSR build")
# ----------------------------------------------------------------------------
def group_observations(ids, n_rows):
    """Stable grouping of observation indices by row id.

    Follows ``_build_index_lists`` (poisson_mf_cavi.py:73-84, hpf_cavi.py:97-107,
    gaussian_mf_cavi_bias.py:69-86): observation t is appended to the list of row
    ids[t] while scanning t = 0..nnz-1, so each row keeps original order.  That is a
    stable sort of arange(nnz) by id.  Returns (row_ptr int64[n_rows+1], perm int64[nnz]).
    """
    ids = np.asarray(ids, dtype=np.int64)
    counts = np.bincount(ids, minlength=n_rows)
    row_ptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.cumsum(counts, out=row_ptr[1:])
    perm = np.argsort(ids, kind="stable").astype(np.int64)
    return row_ptr, perm


def infer_dimensions(u, i):
    """poisson_mf_cavi.py:44-46 -- sizes come from the TRAIN ids only."""
    return int(np.max(u)) + 1, int(np.max(i)) + 1


# ----------------------------------------------------------------------------
# a2: initial variational parameters (PCG64 draw order matters)
# ----------------------------------------------------------------------------
def poisson_init(n_users, n_items, K, a0, b0, seed):
    """poisson_mf_cavi.py:50-71: a = a0 + Gamma(1, 0.1) users then items; b = b0."""
    rng = np.random.default_rng(seed)
    a_t = a0 + rng.gamma(1.0, 0.1, size=(n_users, K))
    a_b = a0 + rng.gamma(1.0, 0.1, size=(n_items, K))
    b_t = b0 * np.ones((n_users, K))
    b_b = b0 * np.ones((n_items, K))
    return dict(a_theta=a_t, b_theta=b_t, a_beta=a_b, b_beta=b_b,
                E_theta=a_t / b_t, E_beta=a_b / b_b)


def hpf_init(n_users, n_items, K, cfg, seed):
    """hpf_cavi.py:66-89: draw order a_theta, b_theta, a_beta, b_beta."""
    rng = np.random.default_rng(seed)
    a_t = cfg["a"] + rng.gamma(1.0, 0.1, size=(n_users, K))
    b_t = cfg["b_prime"] + rng.gamma(1.0, 0.1, size=(n_users, K))
    a_b = cfg["c"] + rng.gamma(1.0, 0.1, size=(n_items, K))
    b_b = cfg["d_prime"] + rng.gamma(1.0, 0.1, size=(n_items, K))
    a_xi = cfg["a_prime"] + K * cfg["a"]
    a_eta = cfg["c_prime"] + K * cfg["c"]
    b_xi = cfg["b_prime"] * np.ones(n_users)
    b_eta = cfg["d_prime"] * np.ones(n_items)
    return dict(gamma_a_theta=a_t, gamma_b_theta=b_t, gamma_a_beta=a_b, gamma_b_beta=b_b,
                gamma_a_xi=a_xi, gamma_b_xi=b_xi, gamma_a_eta=a_eta, gamma_b_eta=b_eta,
                E_theta=a_t / b_t, E_beta=a_b / b_b, E_xi=a_xi / b_xi, E_eta=a_eta / b_eta)


def gauss_init(n_users, n_items, K, seed):
    """gaussian_mf_cavi_bias.py:52-67: m = 0.1*N(0,1) users then items; V = I; biases 0."""
    rng = np.random.default_rng(seed)
    m_t = 0.1 * rng.standard_normal((n_users, K))
    m_b = 0.1 * rng.standard_normal((n_items, K))
    eye = np.eye(K)
    return dict(m_theta=m_t, m_beta=m_b,
                V_theta=np.tile(eye[None], (n_users, 1, 1)),
                V_beta=np.tile(eye[None], (n_items, 1, 1)),
                m_user_bias=np.zeros(n_users), m_item_bias=np.zeros(n_items))


# ----------------------------------------------------------------------------
# a3 / a4: Gamma-Poisson row pass (shared by Poisson MF and HPF-CAVI)
# ----------------------------------------------------------------------------
def gamma_row_pass(row_ptr, perm, other_ids, x, E_self, E_oth, shape_prior, rate_prior):
    """One Jacobi pass over the rows of one side.

    Follows poisson_mf_cavi.py:135-164 (users) / :173-194 (items) and
    hpf_cavi.py:126-151 / :162-185.  For row r with observations t (original order):
      rate_t  = max(E_oth[c_t] . E_self[r], 1e-10)
      alloc_t = (x_t / rate_t) * E_oth[c_t] * E_self[r]
      shp[r]  = shape_prior + sum_t alloc_t ;  rte[r] = rate_prior[r] + sum_t E_oth[c_t]
    Empty rows fall back to (shape_prior, rate_prior[r]).  ``rate_prior`` is a vector
    (b0 broadcast for Poisson MF; E_xi / E_eta for HPF).  Reads only OLD E_self.
    """
    R, K = E_self.shape
    shp = np.empty((R, K))
    rte = np.empty((R, K))
    for r in range(R):
        obs = perm[row_ptr[r]:row_ptr[r + 1]]
        if obs.size == 0:
            shp[r] = shape_prior
            rte[r] = rate_prior[r]
            continue
        sub = E_oth[other_ids[obs]]
        own = E_self[r]
        rate = sub @ own
        rate[rate < RATE_FLOOR] = RATE_FLOOR
        alloc = (x[obs][:, None] / rate[:, None]) * sub * own[None, :]
        shp[r] = shape_prior + np.sum(alloc, axis=0)
        rte[r] = rate_prior[r] + np.sum(sub, axis=0)
    return shp, rte


def poisson_sweeps(u, i, x, K, a0, b0, n_sweeps, seed, n_users=None, n_items=None):
    """T full CAVI sweeps of Poisson MF (poisson_mf_cavi.py:86-197), no validation."""
    u = np.asarray(u, dtype=np.int64)
    i = np.asarray(i, dtype=np.int64)
    x = np.asarray(x, dtype=np.float64)
    if n_users is None:
        n_users, n_items = infer_dimensions(u, i)
    st = poisson_init(n_users, n_items, K, a0, b0, seed)
    rp_u, pm_u = group_observations(u, n_users)
    rp_i, pm_i = group_observations(i, n_items)
    b0_u = np.full(n_users, b0)
    b0_i = np.full(n_items, b0)
    for _ in range(n_sweeps):
        st["a_theta"], st["b_theta"] = gamma_row_pass(rp_u, pm_u, i, x, st["E_theta"], st["E_beta"], a0, b0_u)
        st["E_theta"] = st["a_theta"] / st["b_theta"]                      # :167
        st["a_beta"], st["b_beta"] = gamma_row_pass(rp_i, pm_i, u, x, st["E_beta"], st["E_theta"], a0, b0_i)
        st["E_beta"] = st["a_beta"] / st["b_beta"]                         # :197
    st["n_users"], st["n_items"] = n_users, n_items
    return st


# ----------------------------------------------------------------------------
# §8f-4: extended Poisson MF (per-user phi_u and per-item psi_i scalars)
# ----------------------------------------------------------------------------
def poisson_ext_init(n_users, n_items, K, a0, b0, seed):
    """poisson_mf_extended_cavi.py:54-76: draws a_theta, a_beta, a_phi, a_psi in this order; every rate = b0."""
    rng = np.random.default_rng(seed)
    st = dict(a_theta=a0 + rng.gamma(1.0, 0.1, size=(n_users, K)), a_beta=a0 + rng.gamma(1.0, 0.1, size=(n_items, K)),
              a_phi=a0 + rng.gamma(1.0, 0.1, size=n_users), a_psi=a0 + rng.gamma(1.0, 0.1, size=n_items),
              b_theta=b0 * np.ones((n_users, K)), b_beta=b0 * np.ones((n_items, K)),
              b_phi=b0 * np.ones(n_users), b_psi=b0 * np.ones(n_items))
    for f in ("theta", "beta", "phi", "psi"):
        st["E_" + f] = st["a_" + f] / st["b_" + f]
    return st


def poisson_ext_row_pass(row_ptr, perm, other_ids, x, E_self, s_self, E_oth, s_oth, a0, b0):
    """One side of an extended sweep (poisson_mf_extended_cavi.py:110-164 users / :169-216 items).

    Row r, observations t in original order, c_t the other side's id:
      dot_t   = E_oth[c_t] . E_self[r]                       (NOT clamped: the clamped rate_est :137-138 is unused)
      shp[r]  = a0 + sum_t (x_t / dot_t) * E_oth[c_t] * E_self[r]        (:142-143)
      rte[r]  = b0 + sum_t s_oth[c_t] * E_oth[c_t]                        (:147-148)
      E_new   = shp[r] / rte[r]                                           (:160)
      s_shp[r] = a0 + sum_t x_t                                           (:153)
      s_rte[r] = b0 + sum_t s_oth[c_t] * (E_oth[c_t] . E_new)             (:163-164, in-row Gauss-Seidel)
    Empty rows: the shape/rate parameters fall back to their priors but the EXPECTATIONS ARE NOT TOUCHED (:112-118
    `continue`s before :160/:167), so such a row keeps its initial E for ever.  Rows only read their own E_self and the
    other side's tables, so the pass is Jacobi across rows.  Returns shp, rte, s_shp, s_rte, E_new, sE_new.
    """
    R, K = E_self.shape
    shp, rte = np.empty((R, K)), np.empty((R, K))
    s_shp, s_rte = np.empty(R), np.empty(R)
    E_new, sE_new = E_self.copy(), s_self.copy()
    for r in range(R):
        obs = perm[row_ptr[r]:row_ptr[r + 1]]
        if obs.size == 0:
            shp[r], rte[r], s_shp[r], s_rte[r] = a0, b0, a0, b0
            continue
        sub = E_oth[other_ids[obs]]
        sc = s_oth[other_ids[obs]]
        own = E_self[r]
        dot = sub @ own
        with np.errstate(divide="ignore", invalid="ignore"):
            alloc = (x[obs][:, None] / dot[:, None]) * sub * own[None, :]
        shp[r] = a0 + np.sum(alloc, axis=0)
        rte[r] = b0 + np.sum(sub * sc[:, None], axis=0)
        new = shp[r] / rte[r]
        s_shp[r] = a0 + np.sum(x[obs])
        s_rte[r] = b0 + np.sum(sc * (sub @ new))
        E_new[r], sE_new[r] = new, s_shp[r] / s_rte[r]
    return shp, rte, s_shp, s_rte, E_new, sE_new


def poisson_ext_sweeps(u, i, x, K, a0, b0, n_sweeps, seed, n_users=None, n_items=None):
    """T sweeps of the extended model (poisson_mf_extended_cavi.py:89-216), no validation."""
    u = np.asarray(u, dtype=np.int64)
    i = np.asarray(i, dtype=np.int64)
    x = np.asarray(x, dtype=np.float64)
    if n_users is None:
        n_users, n_items = infer_dimensions(u, i)
    st = poisson_ext_init(n_users, n_items, K, a0, b0, seed)
    rp_u, pm_u = group_observations(u, n_users)
    rp_i, pm_i = group_observations(i, n_items)
    for _ in range(n_sweeps):
        st["a_theta"], st["b_theta"], st["a_phi"], st["b_phi"], st["E_theta"], st["E_phi"] = poisson_ext_row_pass(
            rp_u, pm_u, i, x, st["E_theta"], st["E_phi"], st["E_beta"], st["E_psi"], a0, b0)
        st["a_beta"], st["b_beta"], st["a_psi"], st["b_psi"], st["E_beta"], st["E_psi"] = poisson_ext_row_pass(
            rp_i, pm_i, u, x, st["E_beta"], st["E_psi"], st["E_theta"], st["E_phi"], a0, b0)
    st["n_users"], st["n_items"] = n_users, n_items
    return st


def poisson_ext_predict(user_ids, item_ids, st):
    """phi_u * psi_i * (theta_u . beta_i), 0 for unseen ids (poisson_mf_extended_cavi.py:239-258)."""
    user_ids = np.asarray(user_ids, dtype=np.int64)
    item_ids = np.asarray(item_ids, dtype=np.int64)
    out = np.zeros(len(user_ids))
    ok = (user_ids < st["E_theta"].shape[0]) & (item_ids < st["E_beta"].shape[0])
    uu, ii = user_ids[ok], item_ids[ok]
    out[ok] = st["E_phi"][uu] * st["E_psi"][ii] * np.sum(st["E_theta"][uu] * st["E_beta"][ii], axis=1)
    return out


def hpf_sweeps(u, i, x, K, cfg, n_sweeps, seed, n_users=None, n_items=None):
    """T full sweeps of observed-only HPF-CAVI (hpf_cavi.py:109-193), no validation.

    Order per sweep: user pass (old E_theta, E_beta, E_xi) -> E_theta -> xi rate
    (:158) -> item pass (new E_theta, old E_beta, E_eta) -> E_beta -> eta rate (:192).
    """
    u = np.asarray(u, dtype=np.int64)
    i = np.asarray(i, dtype=np.int64)
    x = np.asarray(x, dtype=np.float64)
    if n_users is None:
        n_users, n_items = infer_dimensions(u, i)
    st = hpf_init(n_users, n_items, K, cfg, seed)
    rp_u, pm_u = group_observations(u, n_users)
    rp_i, pm_i = group_observations(i, n_items)
    for _ in range(n_sweeps):
        st["gamma_a_theta"], st["gamma_b_theta"] = gamma_row_pass(
            rp_u, pm_u, i, x, st["E_theta"], st["E_beta"], cfg["a"], st["E_xi"])
        st["E_theta"] = st["gamma_a_theta"] / st["gamma_b_theta"]
        st["gamma_b_xi"] = cfg["b_prime"] + np.sum(st["E_theta"], axis=1)
        st["E_xi"] = st["gamma_a_xi"] / st["gamma_b_xi"]
        st["gamma_a_beta"], st["gamma_b_beta"] = gamma_row_pass(
            rp_i, pm_i, u, x, st["E_beta"], st["E_theta"], cfg["c"], st["E_eta"])
        st["E_beta"] = st["gamma_a_beta"] / st["gamma_b_beta"]
        st["gamma_b_eta"] = cfg["d_prime"] + np.sum(st["E_beta"], axis=1)
        st["E_eta"] = st["gamma_a_eta"] / st["gamma_b_eta"]
    st["n_users"], st["n_items"] = n_users, n_items
    return st


# ----------------------------------------------------------------------------
# a5: Gaussian MF passes
# ----------------------------------------------------------------------------
def gauss_factor_pass(row_ptr, perm, other_ids, x, m_self, V_self, m_oth, V_oth,
                      b_self, b_oth, sigma2, eta2):
    """gaussian_mf_cavi_bias.py:132-165 (users) / :170-201 (items).

    S = sum_t (V_oth[c_t] + m_t m_t^T); V[r] = inv(I/eta2 + S/sigma2);
    m[r] = V[r] @ (sum_t (x_t - b_self[r] - b_oth[c_t]) m_t) / sigma2.
    Rows without observations keep their state (:134-135).  In place on copies.
    """
    R, K = m_self.shape
    m_new = m_self.copy()
    V_new = V_self.copy()
    eye = np.eye(K)
    for r in range(R):
        obs = perm[row_ptr[r]:row_ptr[r + 1]]
        if obs.size == 0:
            continue
        c = other_ids[obs]
        res = x[obs] - b_self[r] - b_oth[c]
        mo = m_oth[c]
        S = (V_oth[c] + np.einsum("nk,nl->nkl", mo, mo)).sum(axis=0)
        V = np.linalg.inv(eye / eta2 + S / sigma2)
        m_new[r] = (1.0 / sigma2) * V @ (mo * res[:, None]).sum(axis=0)
        V_new[r] = V
    return m_new, V_new


def gauss_bias_pass(row_ptr, perm, other_ids, x, m_self, m_oth, b_self, b_oth, sigma2, eta_b2):
    """gaussian_mf_cavi_bias.py:206-232 (users) / :237-263 (items)."""
    b_new = b_self.copy()
    for r in range(m_self.shape[0]):
        obs = perm[row_ptr[r]:row_ptr[r + 1]]
        if obs.size == 0:
            continue
        c = other_ids[obs]
        res = x[obs] - b_oth[c] - m_oth[c] @ m_self[r]
        var = 1.0 / (1.0 / eta_b2 + obs.size / sigma2)
        b_new[r] = (var / sigma2) * res.sum()
    return b_new


def gauss_sweeps(u, i, x, K, sigma2, eta_theta2, eta_beta2, eta_bias2, n_sweeps, seed,
                 bias=True, n_users=None, n_items=None):
    """T sweeps of Gaussian MF CAVI.

    bias=True  -> gaussian_mf_cavi_bias.py:125-263 (4 passes: theta, beta, b_u, b_i)
    bias=False -> gaussian_mf_cavi.py:114-178 (2 passes; biases identically 0)
    """
    u = np.asarray(u, dtype=np.int64)
    i = np.asarray(i, dtype=np.int64)
    x = np.asarray(x, dtype=np.float64)
    if n_users is None:
        n_users, n_items = infer_dimensions(u, i)
    st = gauss_init(n_users, n_items, K, seed)
    rp_u, pm_u = group_observations(u, n_users)
    rp_i, pm_i = group_observations(i, n_items)
    for _ in range(n_sweeps):
        st["m_theta"], st["V_theta"] = gauss_factor_pass(
            rp_u, pm_u, i, x, st["m_theta"], st["V_theta"], st["m_beta"], st["V_beta"],
            st["m_user_bias"], st["m_item_bias"], sigma2, eta_theta2)
        st["m_beta"], st["V_beta"] = gauss_factor_pass(
            rp_i, pm_i, u, x, st["m_beta"], st["V_beta"], st["m_theta"], st["V_theta"],
            st["m_item_bias"], st["m_user_bias"], sigma2, eta_beta2)
        if bias:
            st["m_user_bias"] = gauss_bias_pass(
                rp_u, pm_u, i, x, st["m_theta"], st["m_beta"], st["m_user_bias"],
                st["m_item_bias"], sigma2, eta_bias2)
            st["m_item_bias"] = gauss_bias_pass(
                rp_i, pm_i, u, x, st["m_beta"], st["m_theta"], st["m_item_bias"],
                st["m_user_bias"], sigma2, eta_bias2)
    st["n_users"], st["n_items"] = n_users, n_items
    return st


# ----------------------------------------------------------------------------
# a8-a10: predict and metrics
# ----------------------------------------------------------------------------
def predict(user_ids, item_ids, F_user, F_item, b_user=None, b_item=None, global_mean=0.0):
    """poisson_mf_cavi.py:221-241 / gaussian_mf_cavi_bias.py:291-316.

    Out-of-range ids predict 0 (+ global_mean for the Gaussian model).
    """
    user_ids = np.asarray(user_ids, dtype=np.int64)
    item_ids = np.asarray(item_ids, dtype=np.int64)
    out = np.zeros(len(user_ids))
    ok = (user_ids < F_user.shape[0]) & (item_ids < F_item.shape[0])
    uu, ii = user_ids[ok], item_ids[ok]
    val = np.sum(F_user[uu] * F_item[ii], axis=1)
    if b_user is not None:
        val = b_user[uu] + b_item[ii] + val
    out[ok] = val
    return out + global_mean


def rmse(y_true, y_pred):
    """metrics.py:6-10."""
    return float(np.sqrt(np.mean((np.asarray(y_true, float) - y_pred) ** 2)))


def macro_mae(y_true, y_pred):
    """metrics.py:37-51: mean over distinct true labels of the per-label MAE."""
    y_true = np.asarray(y_true, float)
    per = [np.mean(np.abs(y_true[y_true == lab] - y_pred[y_true == lab])) for lab in np.unique(y_true)]
    return float(np.mean(per))


def gauss_eval(u, i, rating_centered, st, global_mean):
    """gaussian_mf_cavi_bias.py:318-347: drops out-of-range rows FIRST, adds the mean back."""
    u = np.asarray(u, dtype=np.int64)
    i = np.asarray(i, dtype=np.int64)
    ok = (u < st["m_theta"].shape[0]) & (i < st["m_beta"].shape[0])
    if not np.any(ok):
        return float("nan"), float("nan")
    y = np.asarray(rating_centered, float)[ok] + global_mean
    p = predict(u[ok], i[ok], st["m_theta"], st["m_beta"], st["m_user_bias"], st["m_item_bias"], global_mean)
    return rmse(y, p), macro_mae(y, p)


def log_gamma(v):
    """lgamma for arrays without scipy at call sites (metrics.py:64 uses scipy gammaln)."""
    from math import lgamma
    return np.vectorize(lgamma, otypes=[float])(v)


def poisson_lpl(u, i, rating, theta, beta, eps=1e-10):
    """metrics.py:53-66: sum x log max(lam, eps) - lam - lgamma(x+1)."""
    lam = np.maximum(np.sum(theta[u] * beta[i], axis=1), eps)
    rating = np.asarray(rating, float)
    return float(np.sum(rating * np.log(lam) - lam - log_gamma(rating + 1.0)))


def gauss_lpl(u, i, rating, theta, beta, sigma):
    """metrics.py:18-35 (note: squares its ``sigma`` argument)."""
    err2 = (np.asarray(rating, float) - np.sum(theta[u] * beta[i], axis=1)) ** 2
    var = sigma ** 2
    return float(np.sum(-0.5 * np.log(2 * np.pi * var) - err2 / (2 * var)))


# ----------------------------------------------------------------------------
# a6 / a7: HPF-MAP (gradient-based "PyTorch HPF") -- analytic restatement
# ----------------------------------------------------------------------------
def softplus(z):
    """torch.nn.functional.softplus with beta=1, threshold=20 (hpf_pytorch.py:50-64)."""
    z = np.asarray(z)
    return np.where(z > 20.0, z, np.log1p(np.exp(np.minimum(z, 20.0))))


def sigmoid(z):
    return 1.0 / (1.0 + np.exp(-np.asarray(z)))


def hpf_map_loss_grads(P, users, items, ratings, user_scale, item_scale, cfg, dtype=np.float64):
    """Loss and dense gradients of HPF_PyTorch.loss (hpf_pytorch.py:71-184).

    P holds the unconstrained parameters theta_u (N,K), beta_u (M,K), xi_u (N), eta_u (M).
      lam_b = max(theta_u . beta_i, 1e-6)                      (:78-80)
      loss  = sum_b lam - r log lam                              (:83)
            + sum_b s_u sum_k(-a log xi_u + xi_u th_uk - (a-1) log th_uk)   (:145-152)
            + sum_b t_i sum_k(-c log eta_i + eta_i be_ik - (c-1) log be_ik) (:158-165)
            + sum_b s_u(-(a'-1) log xi_u + b' xi_u) + sum_b t_i(-(c'-1) log eta_i + d' eta_i)
    Gradients w.r.t. the unconstrained parameters are chained through
    softplus' = sigmoid, accumulated over duplicate ids in the batch.
    """
    th_raw, be_raw, xi_raw, et_raw = (np.asarray(P[k], dtype) for k in ("theta", "beta", "xi", "eta"))
    users = np.asarray(users, np.int64)
    items = np.asarray(items, np.int64)
    r = np.asarray(ratings, dtype)
    a, c = cfg["a"], cfg["c"]
    ap, bp, cp, dp = cfg["a_prime"], cfg["b_prime"], cfg["c_prime"], cfg["d_prime"]
    K = th_raw.shape[1]
    th = softplus(th_raw[users]); be = softplus(be_raw[items])
    xi = softplus(xi_raw[users]); et = softplus(et_raw[items])
    s = np.asarray(user_scale, dtype)[users]; t = np.asarray(item_scale, dtype)[items]
    dot = np.sum(th * be, axis=1)
    lam = np.maximum(dot, 1e-6)
    loss = np.sum(lam - r * np.log(lam))
    loss += np.sum(s * np.sum(-a * np.log(xi)[:, None] + xi[:, None] * th - (a - 1) * np.log(th), axis=1))
    loss += np.sum(t * np.sum(-c * np.log(et)[:, None] + et[:, None] * be - (c - 1) * np.log(be), axis=1))
    loss += np.sum(s * (-(ap - 1) * np.log(xi) + bp * xi))
    loss += np.sum(t * (-(cp - 1) * np.log(et) + dp * et))
    g = np.where(dot >= 1e-6, 1.0 - r / lam, 0.0)           # clamp kills the gradient below 1e-6
    d_th = g[:, None] * be + s[:, None] * (xi[:, None] - (a - 1) / th)
    d_be = g[:, None] * th + t[:, None] * (et[:, None] - (c - 1) / be)
    d_xi = s * (-K * a / xi + np.sum(th, axis=1) - (ap - 1) / xi + bp)
    d_et = t * (-K * c / et + np.sum(be, axis=1) - (cp - 1) / et + dp)
    G = {k: np.zeros_like(np.asarray(P[k], dtype)) for k in ("theta", "beta", "xi", "eta")}
    np.add.at(G["theta"], users, d_th * sigmoid(th_raw[users]))
    np.add.at(G["beta"], items, d_be * sigmoid(be_raw[items]))
    np.add.at(G["xi"], users, d_xi * sigmoid(xi_raw[users]))
    np.add.at(G["eta"], items, d_et * sigmoid(et_raw[items]))
    return float(loss), G


def adam_dense_step(P, G, M, V, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam single-tensor update (torch/optim/adam.py, defaults, wd=0).

    Dense: rows with zero gradient still decay m, v and move (compare_models.py:288,312).
    """
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    for k in P:
        M[k] += (G[k] - M[k]) * (1.0 - beta1)          # exp_avg.lerp_(grad, 1-beta1)
        V[k] *= beta2
        V[k] += (1.0 - beta2) * G[k] * G[k]
        denom = np.sqrt(V[k]) / np.sqrt(bc2) + eps
        P[k] -= (lr / bc1) * (M[k] / denom)


# ----------------------------------------------------------------------------
# a11: textbook HPF extras -- PARITY UNPINNED (no reference code exists)
# ----------------------------------------------------------------------------
def digamma(v):
    """psi(x) by recurrence to x>=6 then the asymptotic series (float64)."""
    v = np.asarray(v, dtype=np.float64).copy()
    out = np.zeros_like(v)
    for _ in range(6):
        small = v < 6.0
        out = np.where(small, out - 1.0 / v, out)
        v = np.where(small, v + 1.0, v)
    inv = 1.0 / v
    inv2 = inv * inv
    out += np.log(v) - 0.5 * inv - inv2 * (1.0 / 12 - inv2 * (1.0 / 120 - inv2 * (1.0 / 252 - inv2 * (1.0 / 240 - inv2 / 132))))
    return out


def hpf_digamma_row_pass(row_ptr, perm, other_ids, x, G_self, G_oth, E_oth, shape_prior, rate_prior):
    """docs/Models.tex:652-664 multinomial step: phi_k ∝ exp(E log th_k + E log be_k).

    G = exp(psi(shp))/rte is the geometric-mean table, so phi_k = G_self_k G_oth_k / sum_k.
    Shape gets sum_t x_t phi_tk; rate keeps the reference code's observed-only sum of
    arithmetic means (hpf_cavi.py:151).  PARITY UNPINNED.
    """
    R, K = G_self.shape
    shp = np.empty((R, K)); rte = np.empty((R, K))
    for r in range(R):
        obs = perm[row_ptr[r]:row_ptr[r + 1]]
        if obs.size == 0:
            shp[r] = shape_prior; rte[r] = rate_prior[r]
            continue
        c = other_ids[obs]
        w = G_oth[c] * G_self[r][None, :]
        z = np.maximum(w.sum(axis=1), RATE_FLOOR)
        shp[r] = shape_prior + np.sum((x[obs] / z)[:, None] * w, axis=0)
        rte[r] = rate_prior[r] + np.sum(E_oth[c], axis=0)
    return shp, rte


def hpf_digamma_sweeps(u, i, x, K, cfg, n_sweeps, seed):
    """HPF-CAVI with the digamma allocation of docs/Models.tex:631-726.  PARITY UNPINNED."""
    u = np.asarray(u, np.int64); i = np.asarray(i, np.int64); x = np.asarray(x, float)
    n_users, n_items = infer_dimensions(u, i)
    st = hpf_init(n_users, n_items, K, cfg, seed)
    rp_u, pm_u = group_observations(u, n_users)
    rp_i, pm_i = group_observations(i, n_items)
    geo = lambda shp, rte: np.exp(digamma(shp)) / rte
    G_t = geo(st["gamma_a_theta"], st["gamma_b_theta"])
    G_b = geo(st["gamma_a_beta"], st["gamma_b_beta"])
    for _ in range(n_sweeps):
        st["gamma_a_theta"], st["gamma_b_theta"] = hpf_digamma_row_pass(
            rp_u, pm_u, i, x, G_t, G_b, st["E_beta"], cfg["a"], st["E_xi"])
        st["E_theta"] = st["gamma_a_theta"] / st["gamma_b_theta"]
        G_t = geo(st["gamma_a_theta"], st["gamma_b_theta"])
        st["gamma_b_xi"] = cfg["b_prime"] + st["E_theta"].sum(axis=1)
        st["E_xi"] = st["gamma_a_xi"] / st["gamma_b_xi"]
        st["gamma_a_beta"], st["gamma_b_beta"] = hpf_digamma_row_pass(
            rp_i, pm_i, u, x, G_b, G_t, st["E_theta"], cfg["c"], st["E_eta"])
        st["E_beta"] = st["gamma_a_beta"] / st["gamma_b_beta"]
        G_b = geo(st["gamma_a_beta"], st["gamma_b_beta"])
        st["gamma_b_eta"] = cfg["d_prime"] + st["E_beta"].sum(axis=1)
        st["E_eta"] = st["gamma_a_eta"] / st["gamma_b_eta"]
    return st


def hpf_elbo(u, i, x, st, cfg):
    """Observed-only HPF evidence lower bound.  PARITY UNPINNED.

    Gopalan, Hofman & Blei (2015) with the rate term restricted to observed pairs to
    match hpf_cavi.py:149-151:
      sum_obs [ x log(sum_k G_th G_be) - lgamma(x+1) - sum_k E_th E_be ]
      + E[log p(theta|xi)] + E[log p(xi)] + E[log p(beta|eta)] + E[log p(eta)] + entropies.
    (With phi at its optimum, sum_k x phi_k (Elog th + Elog be - log phi_k) = x log sum_k G G.)
    """
    u = np.asarray(u, np.int64); i = np.asarray(i, np.int64); x = np.asarray(x, float)
    K = st["E_theta"].shape[1]
    a_t, b_t = st["gamma_a_theta"], st["gamma_b_theta"]
    a_b, b_b = st["gamma_a_beta"], st["gamma_b_beta"]
    a_x, b_x = st["gamma_a_xi"], st["gamma_b_xi"]
    a_e, b_e = st["gamma_a_eta"], st["gamma_b_eta"]
    Elog = lambda shp, rte: digamma(shp) - np.log(rte)
    Eth, Ebe = a_t / b_t, a_b / b_b
    Exi, Eet = a_x / b_x, a_e / b_e
    Lth, Lbe = Elog(a_t, b_t), Elog(a_b, b_b)
    Lxi, Let = Elog(a_x * np.ones_like(b_x), b_x), Elog(a_e * np.ones_like(b_e), b_e)
    z = np.maximum(np.sum(np.exp(Lth[u] + Lbe[i]), axis=1), RATE_FLOOR)
    like = np.sum(x * np.log(z) - log_gamma(x + 1.0) - np.sum(Eth[u] * Ebe[i], axis=1))
    lg = lambda v: log_gamma(np.asarray(v, float))
    a, c = cfg["a"], cfg["c"]
    p_th = np.sum(a * Lxi[:, None] - lg(a) + (a - 1) * Lth - Exi[:, None] * Eth)
    p_be = np.sum(c * Let[:, None] - lg(c) + (c - 1) * Lbe - Eet[:, None] * Ebe)
    ap, bp, cp, dp = cfg["a_prime"], cfg["b_prime"], cfg["c_prime"], cfg["d_prime"]
    p_xi = np.sum(ap * np.log(bp) - lg(ap) + (ap - 1) * Lxi - bp * Exi)
    p_et = np.sum(cp * np.log(dp) - lg(cp) + (cp - 1) * Let - dp * Eet)
    ent = lambda shp, rte: np.sum(shp - np.log(rte) + lg(shp) + (1 - shp) * digamma(shp))
    H = ent(a_t, b_t) + ent(a_b, b_b) + ent(a_x * np.ones_like(b_x), b_x) + ent(a_e * np.ones_like(b_e), b_e)
    return float(like + p_th + p_be + p_xi + p_et + H)


def topn(F_user, F_item, n, user_rows=None):
    """Dense U V^T top-n item indices per user.  PARITY UNPINNED (no reference code).

    Scores are float32 accumulated in k = 0..K-1 order with separate multiply and add;
    ties broken by ascending item index (total order: score desc, index asc).
    """
    Fu = np.asarray(F_user, np.float32)
    Fi = np.asarray(F_item, np.float32)
    if user_rows is not None:
        Fu = Fu[user_rows]
    scores = np.zeros((Fu.shape[0], Fi.shape[0]), np.float32)
    prod = np.empty_like(scores)
    FiT = np.ascontiguousarray(Fi.T)
    for k in range(Fu.shape[1]):
        np.multiply(Fu[:, k:k + 1], FiT[k][None, :], out=prod)           # float32 product, rounded ...
        np.add(scores, prod, out=scores)                                  # ... then rounded float32 add (no FMA)
    M = Fi.shape[0]
    idx = np.empty((Fu.shape[0], n), np.int64)
    for b, row in enumerate(scores):
        nth = np.partition(row, M - n)[M - n]                # n-th largest value: nothing below it can rank in the top n
        cand = np.nonzero(row >= nth)[0]
        idx[b] = cand[np.lexsort((cand, -row[cand]))[:n]]    # score desc, then index asc (== lexsort of the whole row)
    return idx.astype(np.int32), np.take_along_axis(scores, idx, axis=1)
