"""a11 on the GPU (PARITY UNPINNED: the reference has no code for these; the checker is the oracle's own
restatement of docs/Models.tex:631-726): digamma allocation sweeps and the ELBO."""
import numpy as np
import pytest

from conftest import rel_max
from oracle import pmf_oracle as O

pytestmark = pytest.mark.gpu
HP = dict(a=0.3, a_prime=5.0, b_prime=5.0, c=0.3, c_prime=5.0, d_prime=5.0)


def data(N=400, M=300, nnz=8000, seed=21):
    from prob_matrix_factorization_b200 import synth
    u, i, x = synth.make_ratings(N, M, nnz, seed)
    return u, i, x + 1.0, N, M


@pytest.mark.parametrize("K,seg_len", [(6, 8), (50, 64), (100, 16)])
def test_elbo_matches_restatement(K, seg_len):
    from prob_matrix_factorization_b200.hpf_cavi import HPF_CAVI, HPF_CAVI_Config
    u, i, x, N, M = data()
    T = 4
    m = HPF_CAVI(HPF_CAVI_Config(n_factors=K, max_iter=T, tol=None, verbose=False, **HP), seg_len=seg_len, track_elbo=True)
    m.n_users, m.n_items = N, M
    m.fit_arrays(u, i, x)
    hist = []
    for t in range(1, T + 1):
        st = O.hpf_sweeps(u, i, x, K, HP, t, 42, N, M)
        hist.append(O.hpf_elbo(u, i, x, st, HP))
    assert rel_max(np.array(m.elbo_history_), np.array(hist)) < 1e-5
    total, parts = m.elbo(return_parts=True)
    assert abs(total - hist[-1]) < 1e-5 * abs(hist[-1]) and parts.shape == (6,)


def test_elbo_of_initial_state():
    from prob_matrix_factorization_b200.hpf_cavi import HPF_CAVI, HPF_CAVI_Config
    u, i, x, N, M = data()
    m = HPF_CAVI(HPF_CAVI_Config(n_factors=8, max_iter=0, tol=None, verbose=False, **HP))
    m.n_users, m.n_items = N, M
    m.fit_arrays(u, i, x)
    st = O.hpf_init(N, M, 8, HP, 42)
    ref = O.hpf_elbo(u, i, x, st, HP)
    assert abs(m.elbo() - ref) < 1e-5 * abs(ref)


@pytest.mark.parametrize("K,seg_len", [(6, 8), (50, 64)])
def test_digamma_sweeps_match_restatement_and_raise_elbo(K, seg_len):
    from prob_matrix_factorization_b200.hpf_cavi import HPF_CAVI, HPF_CAVI_Config
    u, i, x, N, M = data()
    T = 6
    m = HPF_CAVI(HPF_CAVI_Config(n_factors=K, max_iter=T, tol=None, verbose=False, **HP), seg_len=seg_len,
                 allocation="digamma", track_elbo=True)
    m.n_users, m.n_items = N, M
    m.fit_arrays(u, i, x)
    st = O.hpf_digamma_sweeps(u, i, x, K, HP, T, 42)
    for k in ("gamma_a_theta", "gamma_b_theta", "gamma_a_beta", "gamma_b_beta", "gamma_b_xi", "gamma_b_eta", "E_theta", "E_beta"):
        assert rel_max(getattr(m, k), st[k]) < 2e-5, k
    ref = O.hpf_elbo(u, i, x, st, HP)
    assert abs(m.elbo_history_[-1] - ref) < 1e-5 * abs(ref)
