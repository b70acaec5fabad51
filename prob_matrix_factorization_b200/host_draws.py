"""Host-side initial draws, bit-identical to the reference's NumPy stream but multi-threaded.

The reference initialises with ``rng = np.random.default_rng(seed)`` and ``a + rng.gamma(1.0, 0.1, size=(R, K))``
(poisson_mf_cavi.py:62-63, hpf_cavi.py:71-80).  ``gamma_shape1`` returns exactly that array and leaves ``rng`` exactly
where NumPy would leave it, using libpmf_b200's parallel replay of the PCG64 + ziggurat stream
(``pmf_numpy_exponential_fill``, csrc/host_draws.cu) for large sizes: 320 M variates at BASELINE config C5 take 3.1 s in
NumPy's single sequential stream.  Small sizes, other bit generators and generators holding a buffered 32-bit draw go
through NumPy itself.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _cabi

MIN_PARALLEL = 1 << 18        # below this NumPy's own loop is as fast as starting threads


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def gamma_shape1(rng, scale, size, offset=0.0, threads=None):
    """``offset + rng.gamma(1.0, scale, size)`` (float64), bit for bit, advancing ``rng`` like NumPy does."""
    shape = (size,) if np.isscalar(size) else tuple(size)
    n = int(np.prod(shape))
    bg = rng.bit_generator
    st = bg.state
    if (n < MIN_PARALLEL or os.environ.get("PMF_HOST_DRAWS", "1") == "0" or st.get("bit_generator") != "PCG64"
            or st.get("has_uint32", 0) != 0):
        return offset + rng.gamma(1.0, scale, size=shape)
    mask = (1 << 64) - 1
    state, inc = int(st["state"]["state"]), int(st["state"]["inc"])
    s_in = (C.c_uint64 * 2)(state >> 64, state & mask)
    i_in = (C.c_uint64 * 2)(inc >> 64, inc & mask)
    s_out = (C.c_uint64 * 2)()
    out = np.empty(shape, dtype=np.float64)
    _cabi.call("pmf_numpy_exponential_fill", s_in, i_in, float(scale), float(offset), n, out.ctypes.data,
               int(threads or host_threads()), s_out)
    st["state"]["state"] = (int(s_out[0]) << 64) | int(s_out[1])
    bg.state = st
    return out


MIN_PARALLEL_CAST = 1 << 20


def ids_to_int32(a, name="id"):
    """Contiguous int64 host ids -> int32 (with the range check of ``ratings.as_id_array``), by all host cores."""
    a = np.ascontiguousarray(a)
    if a.dtype != np.int64 or a.size < MIN_PARALLEL_CAST:
        return None
    out = np.empty(a.shape, dtype=np.int32)
    lo, hi = C.c_int64(0), C.c_int64(0)
    _cabi.call("pmf_host_i64_to_i32", a.ctypes.data, a.size, out.ctypes.data, C.byref(lo), C.byref(hi), host_threads())
    if lo.value < 0 or hi.value > np.iinfo(np.int32).max - 1:
        raise ValueError(f"{name} ids must lie in [0, 2^31-2]")
    return out


def to_float32(a):
    """float64 host array -> float32 (NumPy's rounding), by all host cores for large arrays."""
    a = np.asarray(a)
    if a.dtype != np.float64 or a.size < MIN_PARALLEL_CAST or not a.flags.c_contiguous:
        return np.asarray(a, dtype=np.float32)
    out = np.empty(a.shape, dtype=np.float32)
    _cabi.call("pmf_host_f64_to_f32", a.ctypes.data, a.size, out.ctypes.data, host_threads())
    return out


def divide(a, b):
    """``a / b`` for a float64 array ``a`` and an equally shaped float64 array or a scalar ``b`` (IEEE division, bit-identical
    to NumPy), by all host cores for large arrays."""
    a = np.asarray(a)
    scalar = np.isscalar(b)
    if (a.dtype != np.float64 or a.size < MIN_PARALLEL_CAST or not a.flags.c_contiguous
            or (not scalar and (b.dtype != np.float64 or b.shape != a.shape or not b.flags.c_contiguous))):
        return a / b
    out = np.empty(a.shape, dtype=np.float64)
    _cabi.call("pmf_host_divide_f64", a.ctypes.data, None if scalar else b.ctypes.data, float(b) if scalar else 0.0, a.size,
               out.ctypes.data, host_threads())
    return out
