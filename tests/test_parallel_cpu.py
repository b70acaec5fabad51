"""N>1 host logic on CPU: world_size-2 gloo; the per-shard maths is stood in for by the oracle (tests only).

* the round-2 scheme end to end (test_user_range_sharding_gloo): every rank takes 1/world of the list, a histogram all-reduce
  gives nnz-balanced user ranges, a stable split + all-to-all routes each rating to its user's owner in original order, the
  user pass is local, the item pass's per-rank partial row sums are all-reduced ("sufficient statistics combine") -- equal
  to the single-rank result;
* RowExchange (the end-of-fit gather of owned rows) on row-aligned shards, and the bounds / chunk / ownership helpers."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pmf_oracle as O
from prob_matrix_factorization_b200 import synth
from prob_matrix_factorization_b200.parallel import RowExchange, balanced_row_bounds


def test_balanced_bounds_properties():
    rng = np.random.default_rng(0)
    for n_rows, nnz, parts in [(10, 0, 3), (1, 5, 4), (1000, 20_000, 8), (50, 10_000, 7)]:
        ids = np.minimum((n_rows * rng.random(nnz) ** 2).astype(np.int64), n_rows - 1)
        row_ptr, _ = O.group_observations(ids, n_rows)
        b = balanced_row_bounds(row_ptr, parts)
        assert b[0] == 0 and b[-1] == n_rows and np.all(np.diff(b) >= 0) and len(b) == parts + 1
        if nnz:
            per = np.diff(row_ptr[b])
            assert per.sum() == nnz
            assert per.max() <= nnz / parts + np.diff(row_ptr).max()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N, M, nnz, K = 400, 300, 6000, 5
        u, i, x = synth.make_ratings(N, M, nnz, seed=5)
        u = u.astype(np.int64); i = i.astype(np.int64); x = x.astype(np.float64) + 1.0
        cfg = dict(a=0.3, a_prime=5.0, b_prime=5.0, c=0.3, c_prime=5.0, d_prime=5.0)
        st = O.hpf_init(N, M, K, cfg, 42)
        rp_u, pm_u = O.group_observations(u, N)
        rp_i, pm_i = O.group_observations(i, M)
        xu = RowExchange(balanced_row_bounds(rp_u, world))
        xi = RowExchange(balanced_row_bounds(rp_i, world))
        E_t = torch.from_numpy(st["E_theta"].copy()); E_b = torch.from_numpy(st["E_beta"].copy())
        E_x = torch.from_numpy(st["E_xi"].copy()); E_e = torch.from_numpy(st["E_eta"].copy())

        def shard_pass(rp, pm, other, E_self, E_oth, shape, prior, ex):
            lo, hi = ex.bounds[rank], ex.bounds[rank + 1]
            # only this rank's rows are computed; other rows of the replicated table stay stale until gather
            shp, rte = O.gamma_row_pass(rp[lo:hi + 1], pm, other, x, E_self.numpy()[lo:hi], E_oth.numpy(), shape,
                                        prior.numpy()[lo:hi])
            E_self[lo:hi] = torch.from_numpy(shp / rte)

        for _ in range(3):
            shard_pass(rp_u, pm_u, i, E_t, E_b, cfg["a"], E_x, xu)
            lo, hi = xu.bounds[rank], xu.bounds[rank + 1]
            E_x[lo:hi] = st["gamma_a_xi"] / (cfg["b_prime"] + E_t[lo:hi].sum(1))
            xu.gather(E_t, E_x)
            shard_pass(rp_i, pm_i, u, E_b, E_t, cfg["c"], E_e, xi)
            lo, hi = xi.bounds[rank], xi.bounds[rank + 1]
            E_e[lo:hi] = st["gamma_a_eta"] / (cfg["d_prime"] + E_b[lo:hi].sum(1))
            xi.gather(E_b, E_e)
        ref = O.hpf_sweeps(u, i, x, K, cfg, 3, 42, N, M)
        err = max(np.abs(E_t.numpy() - ref["E_theta"]).max(), np.abs(E_b.numpy() - ref["E_beta"]).max(),
                  np.abs(E_x.numpy() - ref["E_xi"]).max(), np.abs(E_e.numpy() - ref["E_eta"]).max())
        q.put((rank, float(err), xu.bytes_last))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_sweeps_equal_single_rank_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, nbytes in res:
        assert err < 1e-12, (rank, err)      # row-aligned shards: bit-identical maths, no cross-rank sums
        assert nbytes > 0


def _worker_user_sharded(rank, world, port, q):
    from prob_matrix_factorization_b200.parallel import balanced_bounds_from_counts
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N, M, nnz, K, T = 500, 350, 9000, 4, 3
        u, i, x = synth.make_ratings(N, M, nnz, seed=8)
        u = u.astype(np.int64); i = i.astype(np.int64); x = x.astype(np.float64) + 1.0
        cfg = dict(a=0.3, a_prime=5.0, b_prime=5.0, c=0.3, c_prime=5.0, d_prime=5.0)
        # --- routing, as ratings.DeviceRatings._route_to_owner does it on the device ---
        per = (nnz + world - 1) // world
        lo, hi = rank * per, min((rank + 1) * per, nnz)
        uc, ic, xc = u[lo:hi], i[lo:hi], x[lo:hi]                       # this rank's piece of the list
        counts = torch.from_numpy(np.bincount(uc, minlength=N))
        dist.all_reduce(counts)
        bounds = balanced_bounds_from_counts(counts.numpy(), world)
        owner = np.searchsorted(bounds[1:], uc, side="right")
        order = np.argsort(owner, kind="stable")                          # = pmf_coo_partition
        send = torch.from_numpy(np.bincount(owner, minlength=world))
        recv = torch.zeros(world, dtype=torch.int64)
        dist.all_to_all_single(recv, send)
        got = []
        for src in (uc[order], ic[order], xc[order]):
            dst = torch.zeros(int(recv.sum()), dtype=torch.from_numpy(src).dtype)
            dist.all_to_all_single(dst, torch.from_numpy(np.ascontiguousarray(src)), recv.tolist(), send.tolist())
            got.append(dst.numpy())
        ul, il, xl = got
        mine = (u >= bounds[rank]) & (u < bounds[rank + 1])
        routed_ok = np.array_equal(ul, u[mine]) and np.array_equal(il, i[mine]) and np.array_equal(xl, x[mine])   # original order kept
        # --- sweeps: local user pass; item pass = per-rank partial sums, all-reduced, then the row update ---
        st = O.hpf_init(N, M, K, cfg, 42)
        E_t, E_b, E_x, E_e = st["E_theta"].copy(), st["E_beta"].copy(), st["E_xi"].copy(), st["E_eta"].copy()
        ulo, uhi = int(bounds[rank]), int(bounds[rank + 1])
        rp_u, pm_u = O.group_observations(ul - ulo, uhi - ulo)
        rp_i, pm_i = O.group_observations(il, M)
        for _ in range(T):
            shp, rte = O.gamma_row_pass(rp_u, pm_u, il, xl, E_t[ulo:uhi], E_b, cfg["a"], E_x[ulo:uhi])
            E_t[ulo:uhi] = shp / rte
            E_x[ulo:uhi] = st["gamma_a_xi"] / (cfg["b_prime"] + E_t[ulo:uhi].sum(1))
            # partial sums over THIS rank's ratings: zero priors give E_beta * sum (x/rate) E_theta  and  sum E_theta
            part_a, part_b = O.gamma_row_pass(rp_i, pm_i, ul, xl, E_b, E_t, 0.0, np.zeros(M))
            stats = torch.from_numpy(np.concatenate([part_a, part_b], axis=1))
            dist.all_reduce(stats)                                         # the sufficient-statistics combine
            tot = stats.numpy()
            shp_b, rte_b = cfg["c"] + tot[:, :K], E_e[:, None] + tot[:, K:]
            E_b = shp_b / rte_b
            E_e = st["gamma_a_eta"] / (cfg["d_prime"] + E_b.sum(1))
        # owned user rows of every rank together = the full table
        full_t = torch.zeros(N, K, dtype=torch.float64); full_t[ulo:uhi] = torch.from_numpy(E_t[ulo:uhi])
        dist.all_reduce(full_t)
        ref = O.hpf_sweeps(u, i, x, K, cfg, T, 42, N, M)
        err = max(np.abs(full_t.numpy() - ref["E_theta"]).max(), np.abs(E_b - ref["E_beta"]).max(), np.abs(E_e - ref["E_eta"]).max())
        q.put((rank, bool(routed_ok), float(err)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_user_range_sharding_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_user_sharded, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, routed_ok, err in res:
        assert routed_ok, rank                  # every rank holds exactly its users' ratings, in original order
        assert err < 1e-12, (rank, err)          # cross-rank sums change only the float64 summation order


def test_item_chunks_and_owned_ranges_partition_the_items():
    """Multi-GPU item pass: chunks (processing order) x ranks tile [0, n_items) exactly once."""
    from prob_matrix_factorization_b200.parallel import item_chunk_bounds, owned_item_ranges
    for n_items, chunks, world in [(500_000, 4, 8), (10, 4, 3), (3, 4, 2), (7, 1, 8), (12_000, 3, 4), (5, 8, 8)]:
        be = item_chunk_bounds(n_items, chunks, "equal")
        assert max(np.diff(be)) - min(np.diff(be)) <= 1 and be[0] == 0 and be[-1] == n_items
        b = item_chunk_bounds(n_items, chunks, "shrink")
        assert b[0] == 0 and b[-1] == n_items and all(x < y for x, y in zip(b, b[1:]))
        sizes = [y - x for x, y in zip(b, b[1:])]
        assert sizes == sorted(sizes, reverse=True) or n_items < 20          # the exposed (last) chunk is the smallest
        seen = np.zeros(n_items, dtype=np.int32)
        for r in range(world):
            ranges = owned_item_ranges(n_items, chunks, world, r, "shrink")
            assert len(ranges) == len(b) - 1
            for c, (lo, hi) in enumerate(ranges):
                assert b[c] <= lo <= hi <= b[c + 1]
                seen[lo:hi] += 1
        assert (seen == 1).all()


def test_balanced_bounds_from_counts():
    from prob_matrix_factorization_b200.parallel import balanced_bounds_from_counts, balanced_row_bounds
    rng = np.random.default_rng(0)
    counts = rng.integers(0, 50, size=1000)
    counts[17] = 5000                                              # one heavy row
    for parts in (1, 2, 3, 8):
        b = balanced_bounds_from_counts(counts, parts)
        row_ptr = np.concatenate([[0], np.cumsum(counts)])
        assert np.array_equal(b, balanced_row_bounds(row_ptr, parts))
        assert b[0] == 0 and b[-1] == 1000 and np.all(np.diff(b) >= 0)
        per = np.array([counts[b[p]:b[p + 1]].sum() for p in range(parts)])
        assert per.sum() == counts.sum() and per.max() <= counts.sum() / parts + counts.max()
