"""a1 on the GPU: device grouping is bit-exact against the oracle's stable sort (= the
reference's _build_index_lists, pinned in test_oracle_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import pmf_oracle as O

pytestmark = pytest.mark.gpu


def _build(key, other, val, n_rows, seg_len=64):
    from prob_matrix_factorization_b200.ratings import Grouped
    dev = torch.device("cuda", 0)
    k = torch.from_numpy(key.astype(np.int32)).to(dev)
    o = torch.from_numpy(other.astype(np.int32)).to(dev)
    v = torch.from_numpy(val.astype(np.float32)).to(dev)
    return Grouped.build(k, o, v, n_rows, seg_len)


def _check(g, key, other, val, n_rows):
    row_ptr, perm = O.group_observations(key, n_rows)
    assert np.array_equal(g.row_ptr().astype(np.int64), row_ptr)
    assert np.array_equal(g.perm().astype(np.int64), perm)
    assert np.array_equal(g.col(), other[perm].astype(np.int32))
    assert np.array_equal(g.val(), val[perm].astype(np.float32))


@pytest.mark.parametrize("n_rows,nnz,seed", [
    (1, 1, 0),                 # single observation
    (7, 5, 1),                 # fewer observations than rows -> empty rows
    (300, 4096, 2),            # exactly one sort tile
    (300, 4097, 3),            # one element into a second tile
    (70_000, 250_001, 4),      # 17-bit keys -> 3 radix passes, ragged tail
    (2_000_000, 1_000_003, 5), # 21-bit keys, mostly empty rows
    (3, 100_000, 6),           # three giant rows (many segments per row)
])
def test_grouping_bit_exact(n_rows, nnz, seed):
    rng = np.random.default_rng(seed)
    key = np.minimum((n_rows * rng.random(nnz) ** 2.0).astype(np.int64), n_rows - 1)
    other = rng.integers(0, 1000, nnz)
    val = rng.integers(0, 6, nnz).astype(np.float32)
    g = _build(key, other, val, n_rows)
    assert g.nnz == nnz and g.n_rows == n_rows
    _check(g, key, other, val, n_rows)
    # segment accounting: every row contributes max(1, ceil(len/seg_len)) segments
    lens = np.bincount(key, minlength=n_rows)
    assert g.n_segments == int(np.maximum(1, -(-lens // g.seg_len)).sum())
    assert g.n_multi_rows == int((lens > g.seg_len).sum())
    g.free()


def test_golden_lists(golden):
    g = golden("poisson")
    for key, other, n, pk, ck in ((g["u"], g["i"], g["n_users"], "user_perm", "user_counts"),
                                  (g["i"], g["u"], g["n_items"], "item_perm", "item_counts")):
        grp = _build(key, other, g["x"], n)
        assert np.array_equal(grp.perm().astype(np.int64), g[pk])
        assert np.array_equal(np.diff(grp.row_ptr()), g[ck])


def test_empty_and_invalid():
    from prob_matrix_factorization_b200 import _cabi
    z = np.zeros(0, dtype=np.int64)
    g = _build(z, z, z.astype(np.float32), 5)
    assert g.nnz == 0 and g.n_segments == 5 and np.array_equal(g.row_ptr(), np.zeros(6, np.int32))
    with pytest.raises(_cabi.PMFError, match="out of range"):
        _build(np.array([0, 9]), np.array([0, 0]), np.ones(2, np.float32), 5)


def test_partition_and_slice():
    from prob_matrix_factorization_b200.parallel import balanced_row_bounds
    rng = np.random.default_rng(11)
    n_rows, nnz = 5000, 200_000
    key = np.minimum((n_rows * rng.random(nnz) ** 2.5).astype(np.int64), n_rows - 1)
    other = rng.integers(0, 777, nnz)
    val = rng.random(nnz).astype(np.float32)
    g = _build(key, other, val, n_rows)
    row_ptr, perm = O.group_observations(key, n_rows)
    for parts in (1, 2, 3, 8):
        b = g.partition(parts)
        assert np.array_equal(b, balanced_row_bounds(row_ptr, parts))
        assert b[0] == 0 and b[-1] == n_rows and np.all(np.diff(b) >= 0)
        sizes = np.diff(row_ptr[b])
        assert sizes.max() - nnz / parts <= np.diff(row_ptr).max()  # off by at most one row
    b = g.partition(3)
    for r in range(3):
        s = g.slice(int(b[r]), int(b[r + 1]))
        lo, hi = row_ptr[b[r]], row_ptr[b[r + 1]]
        assert s.row_offset == b[r] and s.nnz == hi - lo
        assert np.array_equal(s.row_ptr().astype(np.int64), row_ptr[b[r]:b[r + 1] + 1] - lo)
        assert np.array_equal(s.perm().astype(np.int64), perm[lo:hi])
        assert np.array_equal(s.col(), other[perm[lo:hi]].astype(np.int32))
        s.free()
