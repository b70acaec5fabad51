"""A few sweeps + ELBO at C3 (for ncu launch lists)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from prob_matrix_factorization_b200 import synth
from prob_matrix_factorization_b200.hpf_cavi import HPF_CAVI, HPF_CAVI_Config
w, (u, i, x) = synth.workload_ratings("c3")
hp = dict(a=0.3, a_prime=5.0, b_prime=5.0, c=0.3, c_prime=5.0, d_prime=5.0)
m = HPF_CAVI(HPF_CAVI_Config(n_factors=w.n_factors, max_iter=3, tol=None, verbose=False, **hp), track_elbo=True)
m.n_users, m.n_items = w.n_users, w.n_items
m.fit_arrays(u, i, x + np.float32(1.0))
torch.cuda.synchronize()
print("elbo history", m.elbo_history_)
