"""a2: the multi-threaded host replay of NumPy's PCG64 + ziggurat gamma(1.0, scale) stream (pmf_numpy_exponential_fill) is
bit-identical to the installed NumPy -- values AND the generator state afterwards -- for every thread count, size and
seed, including sizes that are not multiples of the block size and consecutive draws from one generator (the reference
draws a_theta, b_theta, a_beta, b_beta back to back: hpf_cavi.py:71-80)."""
import numpy as np
import pytest

from prob_matrix_factorization_b200 import host_draws


@pytest.mark.parametrize("seed,n,threads", [(42, 300_001, 1), (42, 300_001, 5), (7, 2_000_003, 8), (123456, 1 << 20, 3),
                                            (0, 262_144, 16), (99, 5_000_000, 7)])
def test_gamma_shape1_is_bit_identical_to_numpy(seed, n, threads):
    ref_rng = np.random.default_rng(seed)
    want = 0.3 + ref_rng.gamma(1.0, 0.1, size=n)
    rng = np.random.default_rng(seed)
    got = host_draws.gamma_shape1(rng, 0.1, n, offset=0.3, threads=threads)
    assert got.dtype == np.float64 and np.array_equal(got, want)
    assert rng.bit_generator.state == ref_rng.bit_generator.state
    assert np.array_equal(rng.standard_normal(5), ref_rng.standard_normal(5))        # the generators go on identically


def test_consecutive_draws_and_2d_shapes_like_the_models():
    seed, N, M, K = 42, 9000, 4000, 33
    ref = np.random.default_rng(seed)
    want = [0.3 + ref.gamma(1.0, 0.1, size=(N, K)), 5.0 + ref.gamma(1.0, 0.1, size=(N, K)),
            0.3 + ref.gamma(1.0, 0.1, size=(M, K)), 5.0 + ref.gamma(1.0, 0.1, size=(M, K))]
    rng = np.random.default_rng(seed)
    got = [host_draws.gamma_shape1(rng, 0.1, (N, K), 0.3), host_draws.gamma_shape1(rng, 0.1, (N, K), 5.0),
           host_draws.gamma_shape1(rng, 0.1, (M, K), 0.3), host_draws.gamma_shape1(rng, 0.1, (M, K), 5.0)]
    for a, b in zip(want, got):
        assert a.shape == b.shape and np.array_equal(a, b)
    assert rng.bit_generator.state == ref.bit_generator.state


def test_small_sizes_and_buffered_generators_fall_back_to_numpy():
    ref, rng = np.random.default_rng(3), np.random.default_rng(3)
    assert np.array_equal(host_draws.gamma_shape1(rng, 0.1, (10, 4), 1.0), 1.0 + ref.gamma(1.0, 0.1, size=(10, 4)))
    ref.integers(0, 10, dtype=np.uint32); rng.integers(0, 10, dtype=np.uint32)        # leaves a buffered 32-bit half-draw
    n = host_draws.MIN_PARALLEL + 17
    assert np.array_equal(host_draws.gamma_shape1(rng, 0.1, n), ref.gamma(1.0, 0.1, size=n))
    assert rng.bit_generator.state == ref.bit_generator.state


def test_parallel_host_casts_equal_numpy():
    rng = np.random.default_rng(0)
    ids = rng.integers(0, 2_000_000, size=host_draws.MIN_PARALLEL_CAST + 12345, dtype=np.int64)
    got = host_draws.ids_to_int32(ids)
    assert got.dtype == np.int32 and np.array_equal(got, ids.astype(np.int32))
    assert host_draws.ids_to_int32(ids[:100]) is None and host_draws.ids_to_int32(ids.astype(np.int32)) is None
    bad = ids.copy(); bad[777] = -1
    with pytest.raises(ValueError):
        host_draws.ids_to_int32(bad)
    bad[777] = 2 ** 31 - 1
    with pytest.raises(ValueError):
        host_draws.ids_to_int32(bad)
    x = rng.standard_normal((host_draws.MIN_PARALLEL_CAST // 8 + 3, 8)) * 1e3
    f = host_draws.to_float32(x)
    assert f.dtype == np.float32 and f.shape == x.shape and np.array_equal(f, x.astype(np.float32))
    from prob_matrix_factorization_b200.ratings import as_id_array
    assert np.array_equal(as_id_array(ids, "user"), ids.astype(np.int32))


def test_parallel_divide_equals_numpy():
    rng = np.random.default_rng(1)
    a = rng.gamma(1.0, 0.1, size=(host_draws.MIN_PARALLEL_CAST // 4 + 5, 4)) + 0.3
    b = rng.gamma(1.0, 0.1, size=a.shape) + 5.0
    assert np.array_equal(host_draws.divide(a, b), a / b)
    assert np.array_equal(host_draws.divide(a, 0.5), a / 0.5)
    assert np.array_equal(host_draws.divide(a[:10], b[:10]), a[:10] / b[:10])
