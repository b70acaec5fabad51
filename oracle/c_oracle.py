"""ctypes wrapper of oracle/pmf_oracle.c (TEST INFRASTRUCTURE ONLY; see that file's header)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libpmf_oracle.so")
_lib = None

i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build(force=False):
    src = os.path.join(HERE, "pmf_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.orc_group.argtypes = [i32p, C.c_int64, C.c_int32, i64p, i64p]
        L.orc_poisson_sweeps.argtypes = [i32p, i32p, f64p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_double,
                                         C.c_double, C.c_int32, f64p, f64p, f64p, f64p, f64p, f64p, C.c_int, C.POINTER(C.c_double)]
        L.orc_hpf_sweeps.argtypes = [i32p, i32p, f64p, C.c_int64, C.c_int32, C.c_int32, C.c_int32] + [C.c_double] * 6 + \
                                    [C.c_int32] + [f64p] * 10 + [C.c_int, C.POINTER(C.c_double)]
        L.orc_predict.argtypes = [i64p, i64p, C.c_int64, f64p, C.c_int32, f64p, C.c_int32, C.c_int32, f64p]
        L.orc_gauss_sweeps.argtypes = [i32p, i32p, f64p, C.c_int64, C.c_int32, C.c_int32, C.c_int32] + [C.c_double] * 4 + \
                                      [C.c_int32, C.c_int32] + [f64p] * 6 + [C.c_int, C.POINTER(C.c_double)]
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def max_threads():
    return int(lib().orc_max_threads())


def group(ids, n_rows):
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    row_ptr = np.zeros(n_rows + 1, dtype=np.int64)
    perm = np.zeros(max(len(ids), 1), dtype=np.int64)
    lib().orc_group(ids, len(ids), n_rows, row_ptr, perm)
    return row_ptr, perm[:len(ids)]


def poisson_sweeps(u, i, x, N, M, K, a0, b0, sweeps, E_theta0, E_beta0, threads=0):
    """Returns dict with E_theta, E_beta, a_*, b_* after `sweeps` sweeps from the given initial means."""
    u = np.ascontiguousarray(u, dtype=np.int32); i = np.ascontiguousarray(i, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=np.float64)
    Et = np.array(E_theta0, dtype=np.float64, order="C"); Eb = np.array(E_beta0, dtype=np.float64, order="C")
    at, bt, ab, bb = np.zeros_like(Et), np.zeros_like(Et), np.zeros_like(Eb), np.zeros_like(Eb)
    secs = C.c_double(0.0)
    lib().orc_poisson_sweeps(u, i, x, len(x), N, M, K, a0, b0, sweeps, Et, Eb, at, bt, ab, bb, threads, C.byref(secs))
    return dict(E_theta=Et, E_beta=Eb, a_theta=at, b_theta=bt, a_beta=ab, b_beta=bb, sweep_seconds=secs.value)


def hpf_sweeps(u, i, x, N, M, K, cfg, sweeps, init, threads=0):
    """`init` as produced by oracle.pmf_oracle.hpf_init (E_theta, E_beta, E_xi, E_eta, gamma_a_xi, gamma_a_eta)."""
    u = np.ascontiguousarray(u, dtype=np.int32); i = np.ascontiguousarray(i, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=np.float64)
    Et = np.array(init["E_theta"], dtype=np.float64, order="C"); Eb = np.array(init["E_beta"], dtype=np.float64, order="C")
    Ex = np.array(init["E_xi"], dtype=np.float64, order="C"); Ee = np.array(init["E_eta"], dtype=np.float64, order="C")
    at, bt, ab, bb = np.zeros_like(Et), np.zeros_like(Et), np.zeros_like(Eb), np.zeros_like(Eb)
    bx, be = np.zeros_like(Ex), np.zeros_like(Ee)
    secs = C.c_double(0.0)
    lib().orc_hpf_sweeps(u, i, x, len(x), N, M, K, cfg["a"], cfg["c"], cfg["b_prime"], cfg["d_prime"],
                         float(init["gamma_a_xi"]), float(init["gamma_a_eta"]), sweeps, Et, Eb, Ex, Ee, at, bt, ab, bb,
                         bx, be, threads, C.byref(secs))
    return dict(E_theta=Et, E_beta=Eb, E_xi=Ex, E_eta=Ee, gamma_a_theta=at, gamma_b_theta=bt, gamma_a_beta=ab,
                gamma_b_beta=bb, gamma_b_xi=bx, gamma_b_eta=be, sweep_seconds=secs.value)


def gauss_sweeps(u, i, x, N, M, K, sigma2, eta_theta2, eta_beta2, eta_bias2, sweeps, init, bias=True, threads=0):
    """`init` as produced by oracle.pmf_oracle.gauss_init (m_theta, m_beta, V_theta, V_beta, m_user_bias, m_item_bias)."""
    u = np.ascontiguousarray(u, dtype=np.int32); i = np.ascontiguousarray(i, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=np.float64)
    st = {k: np.array(init[k], dtype=np.float64, order="C") for k in
          ("m_theta", "m_beta", "V_theta", "V_beta", "m_user_bias", "m_item_bias")}
    secs = C.c_double(0.0)
    lib().orc_gauss_sweeps(u, i, x, len(x), N, M, K, sigma2, eta_theta2, eta_beta2, eta_bias2, int(bool(bias)), sweeps,
                           st["m_theta"], st["m_beta"], st["V_theta"], st["V_beta"], st["m_user_bias"], st["m_item_bias"],
                           threads, C.byref(secs))
    st["sweep_seconds"] = secs.value
    return st


def predict(users, items, F_user, F_item):
    users = np.ascontiguousarray(users, dtype=np.int64); items = np.ascontiguousarray(items, dtype=np.int64)
    Fu = np.ascontiguousarray(F_user, dtype=np.float64); Fi = np.ascontiguousarray(F_item, dtype=np.float64)
    out = np.zeros(len(users))
    lib().orc_predict(users, items, len(users), Fu, Fu.shape[0], Fi, Fi.shape[0], Fu.shape[1], out)
    return out
