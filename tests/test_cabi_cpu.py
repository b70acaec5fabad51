"""CPU checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
declared in include/pmf_b200.h; the product refuses to run without a GPU (no CPU fallback)."""
import ctypes
import subprocess

import numpy as np
import pytest
import torch

from prob_matrix_factorization_b200 import _cabi, build


def test_library_builds_and_exports_header_symbols():
    path = build.build()
    lib = ctypes.CDLL(path)
    declared = _cabi.header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in pmf_b200.h but not exported"
    assert set(_cabi._PROTOTYPES) == set(declared), "ctypes prototypes and header disagree"


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", build.LIB_PATH], capture_output=True, text=True).stdout
    archs = {ln.split(".")[-2] for ln in out.splitlines() if ".cubin" in ln}
    assert archs == {"sm_100a"}, archs


def test_pure_host_entry_points():
    lib = _cabi.load()
    assert lib.pmf_version() >= 100
    assert [lib.pmf_row_stride(k) for k in (1, 8, 10, 50, 64, 100)] == [8, 8, 16, 56, 64, 104]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from prob_matrix_factorization_b200.poisson_mf_cavi import PoissonMFCAVI, PoissonMFCAVIConfig
    m = PoissonMFCAVI(PoissonMFCAVIConfig(n_factors=4, max_iter=1, verbose=False))
    u = np.array([0, 1, 2]); i = np.array([0, 1, 1]); x = np.array([1.0, 2.0, 3.0])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.fit_arrays(u, i, x)
    n = ctypes.c_int(0)
    assert _cabi.load().pmf_device_count(ctypes.byref(n)) != 0   # fails loudly, message available
    assert _cabi.load().pmf_last_error()
