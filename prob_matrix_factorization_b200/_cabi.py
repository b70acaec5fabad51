"""ctypes binding of libpmf_b200.so (include/pmf_b200.h).

The product path has no fallback: if the shared library is missing it is built with nvcc, and
if that fails (or a compute entry point reports an error) a RuntimeError is raised -- matching
the reference scripts' ``try/except Exception`` wrappers (train_all_models.py:25-54).
"""
from __future__ import annotations

import ctypes as C
import os
import re
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PMF_B200_LIB") or os.path.join(_PKG, "libpmf_b200.so")   # override: A/B builds of the library
HEADER_PATH = os.path.join(os.path.dirname(_PKG), "include", "pmf_b200.h")

_lock = threading.Lock()
_lib = None

c_i32p = C.POINTER(C.c_int32)
c_f32p = C.POINTER(C.c_float)
c_f64p = C.POINTER(C.c_double)
VP = C.c_void_p


class LazyAdamState(C.Structure):
    """Mirror of ``pmf_lazy_adam`` (include/pmf_b200.h)."""

    _fields_ = ([(n, VP) for n in ("theta", "beta", "xi", "eta", "m_theta", "m_beta", "m_xi", "m_eta", "v_theta", "v_beta",
                                   "v_xi", "v_eta", "g_theta", "g_beta", "g_xi", "g_eta", "last_user", "last_item",
                                   "claim_user", "claim_item", "step_size", "bc2_sqrt")]
                + [("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("tail1", VP), ("tail2", VP), ("pow5", VP),
                   ("n_pow", C.c_int32)])


PMF_EUNSUPPORTED = -4


class PMFError(RuntimeError):
    """A libpmf_b200 entry point returned a non-zero status (``.status``)."""

    status = 0


# name -> (restype, argtypes).  Pointers to device memory are passed as integers (c_void_p).
_PROTOTYPES = {
    "pmf_version": (C.c_int, []),
    "pmf_last_error": (C.c_char_p, []),
    "pmf_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "pmf_row_stride": (C.c_int, [C.c_int]),
    "pmf_tune": (C.c_int, [C.c_char_p, C.c_int]),
    "pmf_copy_to_host": (C.c_int, [VP, VP, C.c_int64, VP]),
    "pmf_csr_build": (C.c_int, [VP, VP, VP, C.c_int64, C.c_int32, C.c_int32, VP, C.POINTER(VP)]),
    "pmf_csr_slice": (C.c_int, [VP, C.c_int32, C.c_int32, VP, C.POINTER(VP)]),
    "pmf_csr_free": (C.c_int, [VP]),
    "pmf_csr_set_row_offset": (C.c_int, [VP, C.c_int32]),
    "pmf_count_keys": (C.c_int, [VP, C.c_int64, C.c_int32, VP, VP]),
    "pmf_coo_partition": (C.c_int, [VP, VP, VP, C.c_int64, C.c_int32, c_i32p, C.c_int32, VP, VP, VP,
                                    C.POINTER(C.c_int64), VP]),
    "pmf_trim": (C.c_int, []),
    "pmf_host_i64_to_i32": (C.c_int, [VP, C.c_int64, VP, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int32]),
    "pmf_host_f64_to_f32": (C.c_int, [VP, C.c_int64, VP, C.c_int32]),
    "pmf_host_divide_f64": (C.c_int, [VP, VP, C.c_double, C.c_int64, VP, C.c_int32]),
    "pmf_numpy_exponential_fill": (C.c_int, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_double, C.c_double, C.c_int64,
                                             VP, C.c_int32, C.POINTER(C.c_uint64)]),
    "pmf_memcpy_async": (C.c_int, [VP, VP, C.c_int64, VP]),
    "pmf_loop_begin": (C.c_int, [VP, C.POINTER(VP)]),
    "pmf_loop_decide": (C.c_int, [VP, VP, C.c_int32, C.c_double, C.c_int32, C.c_int32, VP, VP, VP]),
    "pmf_loop_end": (C.c_int, [VP]),
    "pmf_loop_run": (C.c_int, [VP, VP]),
    "pmf_loop_free": (C.c_int, [VP]),
    "pmf_csr_nnz": (C.c_int64, [VP]),
    "pmf_csr_rows": (C.c_int32, [VP]),
    "pmf_csr_row_offset": (C.c_int32, [VP]),
    "pmf_csr_segments": (C.c_int32, [VP]),
    "pmf_csr_multi_rows": (C.c_int32, [VP]),
    "pmf_csr_seg_len": (C.c_int32, [VP]),
    "pmf_csr_row_ptr": (VP, [VP]),
    "pmf_csr_perm": (VP, [VP]),
    "pmf_csr_col": (VP, [VP]),
    "pmf_csr_val": (VP, [VP]),
    "pmf_csr_device_bytes": (C.c_int64, [VP]),
    "pmf_csr_partition": (C.c_int, [VP, C.c_int32, c_i32p]),
    "pmf_gamma_pass_workspace_bytes": (C.c_int64, [VP, C.c_int32]),
    "pmf_gamma_pass": (C.c_int, [VP, C.c_int32, C.c_int32, VP, VP, VP, VP, C.c_float, C.c_float, VP,
                                 VP, VP, C.c_float, C.c_float, VP, VP]),
    "pmf_gamma_pass_acc": (C.c_int, [VP, C.c_int32, C.c_int32, VP, VP, VP, VP, C.c_float, C.c_float, VP,
                                     VP, VP, C.c_float, C.c_float, VP, VP, C.c_int32, C.c_int32, VP]),
    "pmf_gamma_combine": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, VP, VP, C.c_int32, VP, VP, VP, VP,
                                    C.c_float, C.c_float, VP, VP, VP, C.c_float, C.c_float, VP]),
    "pmf_gamma_combine_staged": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, VP, C.c_int32, VP, C.c_int32, C.c_int64,
                                           C.c_int32, C.c_int32, VP, VP, VP, VP, C.c_float, C.c_float, VP, VP, VP, C.c_float,
                                           C.c_float, VP]),
    "pmf_gamma_pass_ext": (C.c_int, [VP, C.c_int32, C.c_int32, VP, VP, VP, VP, VP, VP, VP, VP, C.c_float, C.c_float, VP, VP]),
    "pmf_scale_rows": (C.c_int, [VP, VP, C.c_int64, C.c_int32, VP, VP]),
    "pmf_gamma_geomean": (C.c_int, [VP, VP, C.c_int64, C.c_int32, C.c_int32, VP, VP]),
    "pmf_gamma_pass_digamma": (C.c_int, [VP, C.c_int32, C.c_int32, VP, VP, VP, VP, VP, VP, C.c_float, C.c_float, VP,
                                         VP, VP, C.c_float, C.c_float, VP, VP]),
    "pmf_hpf_elbo": (C.c_int, [VP, C.c_int32, C.c_int32] + [VP] * 10 + [C.c_int32] * 4 + [C.c_float] * 6 + [VP, VP]),
    "pmf_topn_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int32, C.c_int32]),
    "pmf_topn_workspace_bytes_ex": (C.c_int64, [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "pmf_topn": (C.c_int, [VP, VP, C.c_int64, VP, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, VP, VP, VP,
                           C.c_int64, VP, VP]),
    "pmf_gauss_packed_stride": (C.c_int, [C.c_int]),
    "pmf_gauss_workspace_bytes": (C.c_int64, [VP, C.c_int32]),
    "pmf_gauss_factor_pass": (C.c_int, [VP, C.c_int32, VP, VP, VP, VP, VP, VP, VP, C.c_float, C.c_float, VP, VP]),
    "pmf_gauss_bias_pass": (C.c_int, [VP, C.c_int32, VP, VP, VP, VP, C.c_float, C.c_float, VP, VP]),
    "pmf_gauss_factor_pass_sharded": (C.c_int, [VP, C.c_int32, VP, VP, VP, VP, VP, VP, VP, C.c_float, C.c_float, VP, VP, VP,
                                                C.c_int32, VP]),
    "pmf_gauss_bias_pass_sharded": (C.c_int, [VP, C.c_int32, VP, VP, VP, VP, C.c_float, C.c_float, VP, VP, VP, C.c_int32, VP]),
    "pmf_hpf_map_loss_grad": (C.c_int, [VP, VP, C.c_int32, VP, C.c_int64, VP, VP, VP, VP, VP, VP, C.c_int32, C.c_int32,
                                        C.c_int32] + [C.c_float] * 6 + [VP, VP, VP, VP, VP, VP, VP]),
    "pmf_adam_dense_step": (C.c_int, [VP, VP, VP, VP, C.c_int64] + [C.c_float] * 5 + [VP]),
    "pmf_hpf_map_lazy_epoch": (C.c_int, [C.POINTER(LazyAdamState), VP, VP, C.c_int32, VP, C.c_int64, C.c_int64, C.c_int64, VP, VP,
                                         C.c_int32, C.c_int32, C.c_int32] + [C.c_float] * 6 + [VP, VP, VP]),
    "pmf_hpf_map_lazy_flush": (C.c_int, [C.POINTER(LazyAdamState), C.c_int32, C.c_int32, C.c_int32, C.c_int64, VP]),
    "pmf_hpf_map_predict": (C.c_int, [VP, VP, C.c_int32, C.c_int64, VP, VP, C.c_int32, C.c_int32, C.c_int32, VP, VP]),
    "pmf_predict": (C.c_int, [VP, VP, C.c_int64, VP, C.c_int32, VP, C.c_int32, C.c_int32, C.c_int32,
                              VP, VP, C.c_float, C.c_int32, VP, VP]),
    "pmf_eval_stats_scratch_bytes": (C.c_int64, []),
    "pmf_eval_stats": (C.c_int, [VP, VP, VP, VP, C.c_int32, C.c_int64, VP, C.c_int32, VP, C.c_int32,
                                 C.c_int32, C.c_int32, VP, VP, C.c_float, C.c_int32, VP, VP, VP]),
}


def header_symbols():
    """Every function name declared in include/pmf_b200.h."""
    with open(HEADER_PATH) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(pmf_[a-z0-9_]+)\s*\(", text)))


def load(build_if_missing=True):
    """Load (building on first use) libpmf_b200.so.  Raises RuntimeError when unavailable."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            if not build_if_missing:
                raise RuntimeError(f"{LIB_PATH} is missing; run python -m prob_matrix_factorization_b200.build")
            from . import build as _build
            _build.build()
        try:
            lib = C.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover - depends on the box
            raise RuntimeError(f"cannot load {LIB_PATH}: {e}") from e
        for name, (res, args) in _PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError here = header/library mismatch, fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(status, what=""):
    if status != 0:
        msg = load().pmf_last_error()
        err = PMFError(f"libpmf_b200 {what} failed (status {status}): {msg.decode() if msg else '?'}")
        err.status = int(status)
        raise err


def call(name, *args):
    """Invoke an int-returning entry point and raise PMFError on failure."""
    check(getattr(load(), name)(*args), name)


def ptr(t):
    """Raw device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch

    return torch.cuda.current_stream().cuda_stream


def require_cuda():
    """The product has no CPU path: refuse to run without a CUDA device."""
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("prob_matrix_factorization_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    n = C.c_int(0)
    call("pmf_device_count", C.byref(n))
    if n.value < 1:
        raise RuntimeError("libpmf_b200: no CUDA device visible")
