"""Row-sharded multi-GPU driver: one process per GPU (torchrun), cells sharded by contiguous row ranges.

Where the path exchanges data:
  * PCA: tiny all-reduces of g-long / g x g / g x l reductions inside the solver (lib-owned NCCL
    communicator, bootstrapped from a 128-byte id broadcast through torch.distributed),
  * PCA -> kNN boundary: ONE all-gather of the X_pca row shards (torch.distributed),
  * kNN -> graph boundary: one all-gather of the (idx, dist) row shards,
  * connectivities + Leiden run replicated (deterministic) on every rank: "replicas only".
Everything else is rank-local.  The collective plumbing below is device-agnostic so the N>1 logic is
covered by world_size-2 gloo tests on CPU (tests/test_distributed_gloo.py).
"""
from __future__ import annotations

import numpy as np

TILE = 128  # shard boundaries are multiples of the kNN tile (sb2_knn_l2_f32 requires q0 % 128 == 0)


def shard_bounds(n: int, world: int) -> list[tuple[int, int]]:
    """Contiguous row ranges, all starts multiples of 128, sizes as equal as that allows."""
    tiles = -(-n // TILE)
    base, extra = divmod(tiles, world)
    out, t0 = [], 0
    for r in range(world):
        t1 = t0 + base + (1 if r < extra else 0)
        out.append((min(t0 * TILE, n), min(t1 * TILE, n)))
        t0 = t1
    return out


def allgather_rows(local, bounds, rank: int, group=None):
    """All-gather row shards of unequal height: pad to the tallest shard, gather, drop the padding."""
    import torch
    import torch.distributed as dist

    world = len(bounds)
    rows = [b - a for a, b in bounds]
    assert local.shape[0] == rows[rank]
    hmax = max(rows)
    pad = torch.zeros((hmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: rows[rank]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][: rows[r]] for r in range(world)], dim=0)


def broadcast_bytes(payload: bytes | None, nbytes: int, src: int = 0, device: str = "cpu", group=None) -> bytes:
    import torch
    import torch.distributed as dist

    t = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src=src, group=group)
    return bytes(t.cpu().numpy().tobytes())


def attach_comm(ctx) -> None:
    """Give `ctx` a lib-owned NCCL communicator spanning the torch.distributed world."""
    import ctypes

    import torch.distributed as dist

    from ._abi import check

    world, rank = dist.get_world_size(), dist.get_rank()
    if world == 1:
        return
    buf = (ctypes.c_char * 128)()
    if rank == 0:
        check(ctx.lib.sb2_comm_unique_id(buf))
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    uid = broadcast_bytes(bytes(buf) if rank == 0 else None, 128, 0, dev)
    check(ctx.lib.sb2_comm_init(ctx.handle, world, rank, uid))
    ctx.n_ranks, ctx.rank = world, rank


def pipeline_sharded(ctx, d_indptr, d_indices, d_data, bounds, rank: int, g: int, *, n_pcs: int = 50,
                     n_neighbors: int = 15, solver: int = 1, resolution: float = 1.0, n_iterations: int = -1,
                     seed: int = 0, ops=None, stage_events=None):
    """pca -> neighbors -> leiden on this rank's CSR row shard; returns device tensors + stage info.

    `ops` defaults to the CUDA drivers in `_ops`; tests inject a stand-in to exercise the sharding
    logic on CPU.
    """
    if ops is None:
        from . import _ops as ops
    def mark(name):
        # stage_events: a list the caller passes to get (name, CUDA event) marks between the stages (bench.py's stage_ms)
        if stage_events is not None:
            import torch

            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            stage_events.append((name, ev))

    n_total = bounds[-1][1]
    r0, r1 = bounds[rank]
    mark("start")
    pca = ops.pca_csr_device(ctx, d_indptr, d_indices, d_data, r1 - r0, g, n_pcs, solver=solver, seed=seed,
                             n_total=n_total)
    mark("pca")
    x_all = allgather_rows(pca["X_pca"], bounds, rank) if len(bounds) > 1 else pca["X_pca"]
    mark("allgather_x_pca")
    idx, dist, kinfo = ops.knn_device(ctx, x_all, n_neighbors, q0=r0, n_query=r1 - r0)
    mark("knn")
    if len(bounds) > 1:
        idx = allgather_rows(idx, bounds, rank)
        dist = allgather_rows(dist, bounds, rank)
    mark("allgather_knn")
    indptr, indices, data, _, _ = ops.fuzzy_simplicial_set_device(ctx, idx, dist, n_total, n_neighbors)
    mark("connectivities")
    member, q, nc, linfo = ops.leiden_device(ctx, indptr, indices, data, n_total, resolution=resolution,
                                             n_iterations=n_iterations, seed=seed)
    mark("leiden")
    return dict(X_pca_local=pca["X_pca"], X_pca=x_all, knn_idx=idx, knn_dist=dist, conn=(indptr, indices, data),
                membership=member, modularity=q, n_communities=nc, pca=pca, knn_info=kinfo, leiden_info=linfo)
