"""Throw-away: host<->device copy strategies + cProfile of the e2e API at the current code."""
import sys, time, os
import numpy as np
sys.path.insert(0, ".")
import torch
import scanpy_b200 as sb
from scanpy_b200._synth import synth_scipy
def bench(f, reps=3):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.time(); r = f(); torch.cuda.synchronize(); ts.append(time.time() - t)
    return min(ts), r
a = np.random.rand(130_000_000).astype(np.float32)  # 520 MB
print("H2D pin_memory+cuda :", bench(lambda: torch.from_numpy(a).pin_memory().to("cuda", non_blocking=True))[0])
print("H2D pageable .cuda():", bench(lambda: torch.from_numpy(a).to("cuda"))[0])
def reg():
    t = torch.from_numpy(a)
    rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), t.numel() * 4, 0)
    d = t.to("cuda", non_blocking=True); torch.cuda.synchronize()
    torch.cuda.cudart().cudaHostUnregister(t.data_ptr())
    return d
print("H2D hostRegister    :", bench(reg)[0])
d = torch.from_numpy(a).to("cuda")
print("D2H .cpu()          :", bench(lambda: d.cpu())[0])
def pinned_d2h():
    h = torch.empty(d.shape, dtype=d.dtype, pin_memory=True); h.copy_(d, non_blocking=True); torch.cuda.synchronize(); return h
print("D2H pinned empty    :", bench(pinned_d2h)[0])
del a, d
X, lab = synth_scipy(1_300_000, 2000, device="cuda")
def e2e():
    ad = sb.MiniAnnData(X)
    sb.pp.pca(ad, n_comps=50); sb.pp.neighbors(ad, n_neighbors=15); sb.tl.leiden(ad, flavor="igraph", n_iterations=-1)
    return ad
e2e()
print("e2e:", bench(e2e, 2)[0])
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); e2e(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
