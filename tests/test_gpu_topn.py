"""a11 top-n scoring on the GPU (PARITY UNPINNED: the checker is the oracle's own definition).  Indices must be
bit-exact in both the exact CUDA-core path and the tcgen05 path (which re-scores its candidates exactly)."""
import numpy as np
import pytest

from oracle import pmf_oracle as O

pytestmark = pytest.mark.gpu


def factors(B, M, K, seed, kind="gamma"):
    rng = np.random.default_rng(seed)
    if kind == "gamma":
        return rng.gamma(0.5, 1.0, (B, K)).astype(np.float32), rng.gamma(0.5, 1.0, (M, K)).astype(np.float32)
    return rng.standard_normal((B, K)).astype(np.float32), rng.standard_normal((M, K)).astype(np.float32)


MODES = [False, True, "unfused"]      # exact CUDA cores / tcgen05 fused filter / tcgen05 with the score matrix in HBM


@pytest.mark.parametrize("tensor", MODES)
@pytest.mark.parametrize("B,M,K,n,kind", [(5, 300, 10, 7, "gamma"), (200, 1000, 100, 50, "gamma"), (130, 2500, 64, 50, "normal"),
                                          (64, 129, 50, 50, "gamma"), (300, 5000, 200, 256, "normal"),
                                          (260, 3000, 160, 100, "normal")])      # K = 160: widest fused shape (2 item stages)
def test_topn_matches_oracle(tensor, B, M, K, n, kind):
    from prob_matrix_factorization_b200.scoring import top_n
    Fu, Fi = factors(B, M, K, seed=B + M, kind=kind)
    idx, score, stats = top_n(Fu, Fi, n, tensor_cores=tensor, return_stats=True)
    ref_idx, ref_score = O.topn(Fu, Fi, n)
    assert np.array_equal(idx, ref_idx)
    assert np.array_equal(score, ref_score)          # same float32 chain -> identical bits
    if tensor:
        assert stats["candidates_rescored"] >= B * n and stats["exact_fallback_rows"] == 0


@pytest.mark.parametrize("tensor", MODES)
def test_topn_ties_and_row_subset(tensor):
    """Massive ties (duplicated items, zero rows): order must fall back to ascending item index."""
    from prob_matrix_factorization_b200.scoring import top_n
    Fu, Fi = factors(40, 600, 16, seed=9)
    Fi[100:400] = Fi[100]            # 300 identical items
    Fi[500:] = 0.0
    Fu[3] = 0.0                      # a user with all-zero scores: top-n = items 0..n-1
    rows = np.array([3, 0, 39, 7, 3], dtype=np.int32)
    idx, score = top_n(Fu, Fi, 20, user_rows=rows, tensor_cores=tensor)
    ref_idx, ref_score = O.topn(Fu, Fi, 20, user_rows=rows)
    assert np.array_equal(idx, ref_idx) and np.array_equal(score, ref_score)
    assert np.array_equal(idx[0], np.arange(20))


@pytest.mark.parametrize("kind", ["gamma", "normal"])
def test_topn_fused_three_levels(kind):
    """60k items = 469 item tiles: dense level 0 (16 tiles), a thresholded middle level (to tile 336) and the final level;
    700 rows = three 256-row blocks, the last one partly padding."""
    from prob_matrix_factorization_b200.scoring import top_n
    Fu, Fi = factors(700, 60_000, 100, seed=77, kind=kind)
    idx, score, stats = top_n(Fu, Fi, 50, tensor_cores=True, return_stats=True)
    ref_idx, ref_score = O.topn(Fu, Fi, 50)
    assert np.array_equal(idx, ref_idx) and np.array_equal(score, ref_score)
    assert stats["exact_fallback_rows"] == 0
    assert 700 * 50 <= stats["candidates_rescored"] < 700 * 1000      # the filter keeps a few hundred of 60k per row


def test_topn_fused_overflow_falls_back_to_exact():
    """More tied items than a row's list holds (4096): those rows must come back through exact scoring, unharmed."""
    from prob_matrix_factorization_b200.scoring import top_n
    Fu, Fi = factors(70, 9000, 32, seed=5)
    Fi[1000:7000] = Fi[17]          # 6000 identical items
    Fi[17] *= 1.0
    Fu[5] = 0.0                     # every score ties at 0
    idx, score, stats = top_n(Fu, Fi, 30, tensor_cores=True, return_stats=True)
    ref_idx, ref_score = O.topn(Fu, Fi, 30)
    assert np.array_equal(idx, ref_idx) and np.array_equal(score, ref_score)
    assert stats["exact_fallback_rows"] >= 1
