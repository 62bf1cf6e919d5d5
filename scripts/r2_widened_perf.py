"""Device times of the widened tools at the bench workload (1.3M cells): tl.umap (200 epochs), spectral init, tl.diffmap's
eigensolver, tl.louvain, tl.paga's arc counts, pp.scale, chunked PCA.  CUDA events on the ctx stream; inputs resident."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from ctypes import byref, c_double, c_int32
from scanpy_b200 import _abi, _ops
from scanpy_b200._abi import check, ptr, LeidenInfo
from scanpy_b200._graph_tools import find_ab_params
from scanpy_b200._synth import synth_scipy
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_300_000
X, lab = synth_scipy(n, 2000)
ctx = _abi.default_context()
ip, ix, dat = _ops.csr_to_device(X)
p = _ops.pca_csr_device(ctx, ip, ix, dat, n, 2000, 50, solver=1)
idx, dist, info = _ops.knn_device(ctx, p["X_pca"], 15)
cp, ci, cw, *_ = _ops.fuzzy_simplicial_set_device(ctx, idx, dist, n, 15)
torch.cuda.synchronize()
nnz = int(cp[-1])
print(f"graph: n={n} arcs={nnz}", flush=True)
def timed(name, fn, reps=1):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); out = None
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1)/reps:.1f} ms", flush=True)
    return out
a, b = find_ab_params(1.0, 0.5)
emb = torch.empty((n, 2), dtype=torch.float32, device="cuda")
timed("umap spectral init (eigsh k=3)", lambda: check(ctx.lib.sb2_umap_spectral_init_f32(ctx.handle, n, ptr(cp), ptr(ci), ptr(cw), 2, 0, ptr(emb))))
timed("umap layout 200 epochs", lambda: check(ctx.lib.sb2_umap_layout_f32(ctx.handle, n, ptr(cp), ptr(ci), ptr(cw), 2, 200, a, b, 1.0, 1.0, 5, 0, ptr(emb))))
e = emb.cpu().numpy(); print("  embedding finite:", bool(np.isfinite(e).all()), "std", e.std(0))
from sklearn.metrics import silhouette_score
sub = np.random.default_rng(0).choice(n, 20000, replace=False)
print("  silhouette of planted labels on 20k sampled cells:", round(float(silhouette_score(e[sub], lab[sub])), 3))
ds = torch.empty(n, dtype=torch.float64, device="cuda")
timed("diffmap transition scale", lambda: check(ctx.lib.sb2_transition_scale_f64(ctx.handle, n, ptr(cp), ptr(ci), ptr(cw), 1, ptr(ds))))
t0 = time.time()
ev, vecs, inf = timed("diffmap eigsh k=15", lambda: _ops.eigsh_scaled_device(ctx, cp, ci, cw, n, 15, d_scale=ds, which="LM"))
print("  evals", np.round(ev[::-1][:6], 6), inf)
member = torch.empty(n, dtype=torch.int32, device="cuda"); q = c_double(); nc = c_int32(); li = LeidenInfo()
timed("louvain", lambda: check(ctx.lib.sb2_louvain_csr_f32(ctx.handle, n, ptr(cp), ptr(ci), ptr(cw), 1.0, 0, ptr(member), byref(q), byref(nc), byref(li))))
print(f"  louvain Q {q.value:.5f} communities {nc.value} levels {li.levels}")
dind = torch.arange(0, n * 14 + 1, 14, dtype=torch.int64, device="cuda"); dcol = idx[:, 1:].contiguous().view(-1)
counts = torch.empty((nc.value, nc.value), dtype=torch.int64, device="cuda")
timed("paga arc counts", lambda: check(ctx.lib.sb2_group_arc_counts(ctx.handle, n, ptr(dind), ptr(dcol), ptr(member), nc.value, ptr(counts))), reps=3)
s1 = torch.empty(2000, dtype=torch.float64, device="cuda"); s2 = torch.empty_like(s1)
timed("scale: column stats", lambda: check(ctx.lib.sb2_csr_col_stats_rows_f32(ctx.handle, n, 2000, ptr(ip), ptr(ix), ptr(dat), None, ptr(s1), ptr(s2))), reps=3)
std = torch.ones(2000, dtype=torch.float64, device="cuda"); d2 = dat.clone()
timed("scale: csr in place (zero_center=False)", lambda: check(ctx.lib.sb2_csr_scale_cols_f32(ctx.handle, n, ptr(ip), ptr(ix), ptr(d2), ptr(std), None, 1, 10.0)), reps=3)
