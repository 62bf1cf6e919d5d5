"""Throw-away: host<->device staging costs (alloc vs copy) and a cProfile of the public-API pipeline."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
import scanpy_b200 as sb
from scanpy_b200._synth import synth_scipy
def tm(f, reps=3):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter(); r = f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
    return min(ts)
nb = 520_000_000
a = np.random.rand(nb // 4).astype(np.float32)
print("alloc pinned 520MB        : %.3f" % tm(lambda: torch.empty(nb // 4, dtype=torch.float32, pin_memory=True)))
p = torch.empty(nb // 4, dtype=torch.float32, pin_memory=True)
print("memcpy pageable->pinned   : %.3f" % tm(lambda: p.copy_(torch.from_numpy(a))))
print("pin_memory() (alloc+copy) : %.3f" % tm(lambda: torch.from_numpy(a).pin_memory()))
print("H2D from pinned           : %.3f" % tm(lambda: p.to("cuda", non_blocking=True)))
print("H2D pageable direct       : %.3f" % tm(lambda: torch.from_numpy(a).to("cuda")))
d = p.to("cuda")
print("D2H into pinned           : %.3f" % tm(lambda: p.copy_(d, non_blocking=True)))
b = np.empty_like(a)
print("memcpy pinned->pageable   : %.3f" % tm(lambda: np.copyto(b, p.numpy())))
print("np.empty+touch pageable   : %.3f" % tm(lambda: np.zeros(nb // 4, np.float32)))
cudart = torch.cuda.cudart()
def reg():
    t = torch.from_numpy(a)
    cudart.cudaHostRegister(t.data_ptr(), nb, 0)
    dd = t.to("cuda", non_blocking=True); torch.cuda.synchronize()
    cudart.cudaHostUnregister(t.data_ptr())
print("register+H2D+unregister   : %.3f" % tm(reg))
def chunked(chunk=32 << 20):
    # double-buffered staging through two fixed pinned chunks
    dd = torch.empty(nb // 4, dtype=torch.float32, device="cuda")
    src = torch.from_numpy(a)
    ne = chunk // 4
    evs = [torch.cuda.Event(), torch.cuda.Event()]
    for i, off in enumerate(range(0, nb // 4, ne)):
        s = stage[i & 1]
        if i >= 2: evs[i & 1].synchronize()
        m = min(ne, nb // 4 - off)
        s[:m].copy_(src[off:off + m])
        dd[off:off + m].copy_(s[:m], non_blocking=True)
        evs[i & 1].record()
    return dd
stage = [torch.empty(8 << 20, dtype=torch.float32, pin_memory=True) for _ in range(2)]
print("chunked 32MB double-buffer: %.3f" % tm(chunked))
del a, b, d, p
X, lab = synth_scipy(1_300_000, 2000)
ad = sb.MiniAnnData(X)
def e2e():
    sb.pp.pca(ad, n_comps=50); sb.pp.neighbors(ad, n_neighbors=15); sb.tl.leiden(ad, flavor="igraph", n_iterations=-1)
e2e()
print("e2e: %.3f" % tm(e2e, 2))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); e2e(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(25)
