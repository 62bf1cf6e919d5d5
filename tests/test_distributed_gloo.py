"""N>1 host logic on CPU: world_size-2 gloo process group (no GPU).  The CUDA drivers are replaced by an
oracle-backed stand-in so that only the sharding / collective plumbing of scanpy_b200.distributed is
under test: shard bounds, unequal-shard all-gather, id broadcast, query offsets, result assembly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from scipy import sparse

from scanpy_b200 import distributed as sbd


def test_shard_bounds_properties():
    for n, w in [(1_300_000, 8), (1000, 2), (129, 2), (100, 4), (128 * 7 + 5, 3), (5, 2)]:
        b = sbd.shard_bounds(n, w)
        assert len(b) == w and b[0][0] == 0 and b[-1][1] == n
        for (a0, a1), (b0, b1) in zip(b, b[1:]):
            assert a1 == b0 and a0 <= a1
        assert all(s % 128 == 0 or s == e for s, e in b)  # non-empty kNN query ranges start on a tile boundary
        sizes = [e - s for s, e in b]
        assert max(sizes) - min(sizes) < 256 or n < 128 * w  # one tile of imbalance + the ragged last tile
    assert sbd.shard_bounds(1_300_000, 8)[0] == (0, 162560)


class _OracleOps:
    """CPU stand-in for scanpy_b200._ops *_device drivers (tests only)."""

    def __init__(self, x_full):
        self.x_full = x_full

    def pca_csr_device(self, ctx, indptr, indices, data, n, g, k, *, solver, seed, n_total):
        from oracle import pca as opca

        r0 = ctx["r0"]
        local = sparse.csr_matrix((data.numpy(), indices.numpy(), indptr.numpy()), shape=(n, g))
        assert (local != self.x_full[r0:r0 + n]).nnz == 0  # the rank really holds its own rows
        full = opca.pca_exact_f64(self.x_full, k)           # stands for the all-reduced solver
        return dict(X_pca=torch.from_numpy(full["X_pca"][r0:r0 + n].astype(np.float32)), iterations=1, converged=True)

    def knn_device(self, ctx, x_all, k, *, q0, n_query):
        from oracle import knn as oknn

        assert q0 % 128 == 0
        idx, d = oknn.knn_brute_queries(x_all.numpy(), q0, q0 + n_query, k)
        return torch.from_numpy(idx.astype(np.int32)), torch.from_numpy(d), dict(n_uncertified=0)

    def fuzzy_simplicial_set_device(self, ctx, idx, d, n, k):
        from oracle import fuzzy as ofz

        c, s, r = ofz.fuzzy_simplicial_set(idx.numpy(), d.numpy(), n, k)
        c.sort_indices()
        return (torch.from_numpy(c.indptr.astype(np.int64)), torch.from_numpy(c.indices), torch.from_numpy(c.data),
                torch.from_numpy(s), torch.from_numpy(r))

    def leiden_device(self, ctx, indptr, indices, data, n, *, resolution, n_iterations, seed):
        from oracle import leiden as old

        adj = sparse.csr_matrix((data.numpy(), indices.numpy(), indptr.numpy()), shape=(n, n))
        m, q, _ = old.leiden(adj, resolution=resolution, n_iterations=n_iterations, seed=seed)
        return torch.from_numpy(m), q, int(m.max()) + 1, dict(passes=1, levels=1, moves=0)


def _make_data():
    rs = np.random.RandomState(0)
    n, g = 700, 60
    centers = rs.standard_normal((5, g)) * 3
    lab = rs.randint(0, 5, n)
    dense = np.maximum(centers[lab] + rs.standard_normal((n, g)), 0).astype(np.float32)
    dense[dense < 1.0] = 0
    return sparse.csr_matrix(dense)


def _run_pipeline(x, bounds, rank):
    r0, r1 = bounds[rank]
    loc = x[r0:r1]
    out = sbd.pipeline_sharded(dict(r0=r0), torch.from_numpy(loc.indptr.astype(np.int64)), torch.from_numpy(loc.indices),
                               torch.from_numpy(loc.data), bounds, rank, x.shape[1], n_pcs=10, n_neighbors=8, ops=_OracleOps(x))
    return out


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = _make_data()
        bounds = sbd.shard_bounds(x.shape[0], world)
        # unequal shards: 700 rows -> 384 + 316
        t = torch.arange(bounds[rank][0], bounds[rank][1], dtype=torch.float32)[:, None].repeat(1, 3)
        g = sbd.allgather_rows(t, bounds, rank)
        assert g.shape == (700, 3) and torch.equal(g[:, 0], torch.arange(700, dtype=torch.float32))
        payload = bytes(range(128)) if rank == 0 else None
        assert sbd.broadcast_bytes(payload, 128, 0, "cpu") == bytes(range(128))
        out = _run_pipeline(x, bounds, rank)
        q.put((rank, out["knn_idx"].numpy(), out["membership"].numpy(), out["X_pca"].numpy(), out["modularity"]))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(300)
def test_two_rank_pipeline_equals_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x = _make_data()
    single = _run_pipeline(x, sbd.shard_bounds(x.shape[0], 1), 0)
    for rank, idx, member, xp, qmod in res:
        np.testing.assert_array_equal(idx, single["knn_idx"].numpy())       # identical kNN lists, same row order
        np.testing.assert_array_equal(member, single["membership"].numpy())  # replicated, deterministic Leiden
        np.testing.assert_allclose(xp, single["X_pca"].numpy(), rtol=1e-6)
        assert qmod == single["modularity"]
