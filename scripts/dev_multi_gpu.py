"""Multi-GPU validation (run with torchrun --nproc-per-node N): sharded pipeline == single-GPU pipeline."""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch, torch.distributed as dist
from scanpy_b200 import _abi, _ops, distributed as sbd
from scanpy_b200._synth import synth_scipy
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl")
ctx = _abi.default_context()
sbd.attach_comm(ctx)
n, g = 200_000, 2000
bounds = sbd.shard_bounds(n, world)
r0, r1 = bounds[rank]
x_local, _ = synth_scipy(n, g, device="cuda", row_start=r0, row_stop=r1)
d = _ops.csr_to_device(x_local)
for solver in (1, 0):
    out = sbd.pipeline_sharded(ctx, *d, bounds, rank, g, n_pcs=50, n_neighbors=15, solver=solver, seed=0)
    torch.cuda.synchronize(); dist.barrier()
    t = time.time()
    out = sbd.pipeline_sharded(ctx, *d, bounds, rank, g, n_pcs=50, n_neighbors=15, solver=solver, seed=0)
    torch.cuda.synchronize(); dist.barrier(); dt = time.time() - t
    if rank == 0:
        print(f"world={world} solver={solver} sharded pipeline {dt:.3f}s  ncomm={out['n_communities']} Q={out['modularity']:.6f} pca_it={out['pca']['iterations']}", flush=True)
    # reference: the same matrix on ONE GPU (rank 0 only), compare
    if rank == 0:
        x_full, lab = synth_scipy(n, g, device="cuda")
        ctx1 = _abi.Context(lr)  # no communicator attached
        d1 = _ops.csr_to_device(x_full)
        one = sbd.pipeline_sharded(ctx1, *d1, sbd.shard_bounds(n, 1), 0, g, n_pcs=50, n_neighbors=15, solver=solver, seed=0)
        xa, xb = out["X_pca"].cpu().numpy().astype(np.float64), one["X_pca"].cpu().numpy().astype(np.float64)
        s = np.sign(np.einsum("ij,ij->j", xa, xb)); rel = np.linalg.norm(xa * s - xb, axis=0) / np.linalg.norm(xb, axis=0)
        ia, ib = out["knn_idx"].cpu().numpy(), one["knn_idx"].cpu().numpy()
        same_rows = (np.sort(ia, 1) == np.sort(ib, 1)).all(1).mean()
        from sklearn.metrics import adjusted_rand_score
        ari = adjusted_rand_score(out["membership"].cpu().numpy(), one["membership"].cpu().numpy())
        print(f"   vs single GPU: X_pca rel err max {rel.max():.2e}, kNN rows identical {same_rows:.6f}, Leiden ARI {ari:.5f}, ARI planted {adjusted_rand_score(lab, out['membership'].cpu().numpy()):.5f}", flush=True)
    dist.barrier()
dist.destroy_process_group()
