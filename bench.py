#!/usr/bin/env python
"""bench.py — cells/s of the pca -> neighbors -> leiden hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference CPU arithmetic (rank 0)

Workload (config.workload): synthetic 1.3M cells x 2000 HVGs CSR (~5 % dense), n_pcs=50, k=15 — the
configuration BASELINE.json's metric is quoted on; it fits one B200.  One "step" = one full pass of the
hot path over that matrix.  `value` = cells/s with the CSR already resident in HBM (CUDA events on the
launching stream, barrier + synchronize on both sides, max over ranks); `e2e` = the same pass through
the public scanpy-signature API (sb.pp.pca / sb.pp.neighbors / sb.tl.leiden) from HOST arrays, all
host<->device copies inside the timed region.  N > 1: cells are row-sharded over ranks (strong scaling:
the matrix is fixed), see scanpy_b200/distributed.py.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "cells/sec end-to-end pca->neighbors->leiden, 1.3M x 2k HVG CSR"
UNIT = "cells/s"


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--n-cells", type=int, default=1_300_000)
    p.add_argument("--n-genes", type=int, default=2000)
    p.add_argument("--n-pcs", type=int, default=50)
    p.add_argument("--k", type=int, default=15)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    return p.parse_args()


def workload_config(a, n_gpus):
    return dict(workload=f"synthetic {a.n_cells} cells x {a.n_genes} HVGs CSR (~5% dense, 32 planted clusters), "
                         f"n_pcs={a.n_pcs} k={a.k} (BASELINE.json configs[2]{'/[3]' if n_gpus > 1 else ''})",
                n_cells=a.n_cells, n_genes=a.n_genes, n_pcs=a.n_pcs, n_neighbors=a.k, leiden="resolution=1, n_iterations=-1",
                parallelism=f"row-sharded x{n_gpus}" if n_gpus > 1 else "single GPU",
                l2="inputs (1.0 GB CSR, 260 MB X_pca) exceed the 126 MB L2; no explicit flush")


# ------------------------------------------------------------------------------------------------
# reference CPU arithmetic on a bounded sample (the oracle: sklearn ARPACK PCA and brute kNN are the
# reference's own call sites; fuzzy set / Leiden are the restatements in oracle/)
def cpu_reference_sample(a, sample_rows: int = 50_000, knn_queries: int = 4096):
    """Times the reference path on rows [0, sample_rows) of the workload and extrapolates each stage to
    n_cells with its own complexity: PCA, connectivities, Leiden linear in n; exact brute-force kNN n^2."""
    import torch
    from threadpoolctl import threadpool_limits  # noqa: F401  (BLAS uses all cores by default)

    from oracle import fuzzy as ofz, knn as oknn, leiden as old, pca as opca
    from scanpy_b200._synth import synth_scipy

    n, s = a.n_cells, min(sample_rows, a.n_cells)
    dev = "cuda" if torch.cuda.is_available() else "cpu"  # data generation only
    x, _ = synth_scipy(n, a.n_genes, device=dev, row_stop=s)
    t = time.perf_counter(); p = opca.pca_arpack(x, a.n_pcs); t_pca = time.perf_counter() - t
    xp = p["X_pca"]
    q = min(knn_queries, s)
    t = time.perf_counter(); oknn.knn_brute_queries(xp, 0, q, a.k); t_q = time.perf_counter() - t
    pair_rate = q * s / t_q                      # distance pairs per second on this host
    idx, dist = oknn.knn_brute(xp, a.k)          # untimed: inputs for the graph stages
    t = time.perf_counter(); c, _, _ = ofz.fuzzy_simplicial_set(idx, dist, s, a.k); t_fz = time.perf_counter() - t
    t = time.perf_counter(); old.leiden(c, seed=0); t_ld = time.perf_counter() - t
    scale = n / s
    est = dict(pca=t_pca * scale, knn=n * float(n) / pair_rate, connectivities=t_fz * scale, leiden=t_ld * scale)
    total = sum(est.values())
    return dict(value=n / total, unit=UNIT, cores=os.cpu_count(), kind="port",
                sample=(f"rows [0,{s}) of the workload: sklearn PCA(arpack) {t_pca:.2f}s, sklearn brute kNN "
                        f"{q}x{s} pairs {t_q:.2f}s ({pair_rate:.3g} pairs/s), oracle fuzzy set {t_fz:.2f}s, oracle Leiden "
                        f"{t_ld:.2f}s; extrapolated to {n} cells (PCA/graph/Leiden linear, exact kNN quadratic): "
                        + ", ".join(f"{k}={v:.1f}s" for k, v in est.items())),
                stage_seconds_extrapolated=est)


def run_reference(a, rank, world):
    if rank != 0:
        return
    t0 = time.perf_counter()
    for _ in range(max(a.warmup, 0)):
        cpu_reference_sample(a, sample_rows=8000, knn_queries=1024)  # warm caches / thread pools cheaply
    vals = []
    for _ in range(a.steps):
        vals.append(cpu_reference_sample(a))
    best = max(vals, key=lambda r: r["value"])
    v = float(np.mean([r["value"] for r in vals]))
    ms = 1e3 * a.n_cells / v
    emit(json.dumps(dict(impl="reference", metric=METRIC, value=v, unit=UNIT, n_gpus=a.gpus, steps=a.steps,
                          warmup=a.warmup, ms_per_step=ms, higher_is_better=True, scaling="strong", vs_baseline=None,
                          dtype="f32", data="synthetic", config=workload_config(a, a.gpus),
                          cpu_baseline=dict(best, value=v),
                          e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                          wall_s=time.perf_counter() - t0)))


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu, self.rows, self._stop_ev = gpu_index, [], threading.Event()

    def run(self):
        while not self._stop_ev.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_ev.wait(0.2)

    def stop(self):
        self._stop_ev.set()
        self.join(timeout=5)
        if not self.rows:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for j, nm in enumerate(names) if any(len(r) > 5 + j and r[5 + j].lower().startswith("active") for r in self.rows)]
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=float(self.rows[0][2]) if self.rows[0][2].replace(".", "").isdigit() else None,
                    power_w_max=max((float(r[3]) for r in self.rows if r[3].replace(".", "").isdigit()), default=None),
                    samples=len(self.rows), reasons=reasons)


def run_b200(a, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import scanpy_b200 as sb
    from scanpy_b200 import _abi, _ops, distributed as sbd
    from scanpy_b200._synth import synth_scipy

    torch.cuda.set_device(local_rank)
    ctx = _abi.default_context()
    if world > 1:
        sbd.attach_comm(ctx)
    n, g = a.n_cells, a.n_genes
    bounds = sbd.shard_bounds(n, world)
    r0, r1 = bounds[rank]
    x_local, _ = synth_scipy(n, g, device="cuda", row_start=r0, row_stop=r1)   # this rank's CSR rows (host)
    d_csr = _ops.csr_to_device(x_local)
    torch.cuda.synchronize()

    def step():
        return sbd.pipeline_sharded(ctx, *d_csr, bounds, rank, g, n_pcs=a.n_pcs, n_neighbors=a.k, solver=1, seed=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        out = step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        out = step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = (ctx.launches - launches0) // max(a.steps, 1)
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / a.steps
    value = n / (ms_step / 1e3)

    # ---- e2e through the public API from host arrays (single GPU: every stage copies in and out) ----
    e2e = None
    if not a.no_e2e:
        if world == 1:
            ad = sb.MiniAnnData(x_local)  # the input object exists before the timed region (like a loaded .h5ad)

            def e2e_step():
                sb.pp.pca(ad, n_comps=a.n_pcs)
                sb.pp.neighbors(ad, n_neighbors=a.k)
                sb.tl.leiden(ad, flavor="igraph", n_iterations=-1)
                return ad
            e2e_step()  # warm-up (pinned staging buffers, pools)
            torch.cuda.synchronize()
            _ops.TRANSFER.update(h2d=0, d2h=0)
            t0 = time.perf_counter()
            reps = max(1, min(a.steps, 2))
            for _ in range(reps):
                ad = e2e_step()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / reps
            e2e = dict(value=n / dt, unit=UNIT, h2d_bytes_per_step=_ops.TRANSFER["h2d"] // reps,
                       d2h_bytes_per_step=_ops.TRANSFER["d2h"] // reps, s_per_step=dt,
                       api="sb.pp.pca -> sb.pp.neighbors -> sb.tl.leiden on a host MiniAnnData (scipy CSR in, numpy/scipy/pandas out)")
        else:
            # sharded e2e: host CSR shard -> device, pipeline, membership + X_pca shard back to host
            def e2e_step():
                d = _ops.csr_to_device(x_local)
                o = sbd.pipeline_sharded(ctx, *d, bounds, rank, g, n_pcs=a.n_pcs, n_neighbors=a.k, solver=1, seed=0)
                return _ops._to_host(o["membership"]), _ops._to_host(o["X_pca_local"])
            e2e_step()
            barrier()
            _ops.TRANSFER.update(h2d=0, d2h=0)
            t0 = time.perf_counter()
            e2e_step()
            barrier()
            dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            e2e = dict(value=n / float(dt.item()), unit=UNIT, h2d_bytes_per_step=_ops.TRANSFER["h2d"],
                       d2h_bytes_per_step=_ops.TRANSFER["d2h"], s_per_step=float(dt.item()),
                       api="scanpy_b200.distributed.pipeline_sharded from per-rank host CSR shards (bytes are per rank)")

    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(ROOT / "MEASURED_PEAKS.json"))
    except Exception:
        pass
    ki = out["knn_info"]
    ach = ki["pass1_flops"] / (ki["pass1_ms"] * 1e-3) / 1e12
    issued = ki["pass1_issued_flops"] / (ki["pass1_ms"] * 1e-3) / 1e12
    peak_tensor = peaks.get("bf16_tflops_sustained", 1400.0)
    sm_max = (clocks or {}).get("sm_max_mhz") or 1965.0
    if ki["pass1_tensor"]:
        # DRAM bytes per launch of the dominant kernel from the committed ncu capture of this exact workload
        # (profiles/r1_ncu_metrics_knn_tiered.txt: dram__bytes_read 12.388 GB + dram__bytes_write 0.780 GB)
        traffic = 13.168e9 if (world == 1 and a.n_cells == 1_300_000 and a.n_pcs == 50 and a.k == 15) else None
        roofline = dict(bound="tensor", kernel="knn_pass1_tc_kernel", achieved=ach, peak=peak_tensor, unit="TFLOP/s",
                        frac=ach / peak_tensor, traffic=traffic,
                        traffic_source=("profiles/r1_ncu_metrics_knn_tiered.txt (ncu --set full, main launch; algorithmic operand bytes: "
                                        "0.33 GB of fp16 images)" if traffic else None),
                        peak_source=("MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else "fallback 1.4 PFLOP/s sustained"),
                        launch_ms=ki["pass1_ms"], algorithmic_flops_per_launch=ki["pass1_flops"],
                        issued_tensor_tflops=issued, frac_issued=issued / peak_tensor,
                        note=("achieved counts the ALGORITHMIC 2*n_q*n*d flops of the pairwise-distance sweep over the summed duration "
                              "of its launches (pilot wave + rest; rows the fp16 tier cannot certify are swept again in split "
                              "precision, n_resweep of them). The fp16 tier issues K = d+3 padded to 64 instead of 50, plus a 1/16 "
                              "sample of the candidate tiles for the starting threshold: issued_tensor_tflops / frac_issued say how "
                              "busy the tensor pipe is. Each 256x128 score tile is 8 MMAs (512 tensor cycles) against one TMEM "
                              "hand-off + 4 tcgen05.ld round trips per epilogue warp, so the sweep is paced by the accumulator "
                              "hand-off, not by the MMAs (DESIGN.md section 5)"),
                        n_resweep=ki.get("n_resweep", 0),
                        share_of_step=ki["pass1_ms"] / ms_step)
    else:
        fp32_peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12
        roofline = dict(bound="tensor", kernel="knn_pass1_kernel", achieved=ach, peak=peak_tensor, unit="TFLOP/s",
                        frac=ach / peak_tensor, traffic=None, launch_ms=ki["pass1_ms"], algorithmic_flops_per_launch=ki["pass1_flops"],
                        note=f"fp32 FFMA path (d too large for the tensor-core tiles); CUDA-core ceiling {fp32_peak:.1f} TFLOP/s",
                        frac_fp32_ffma=ach / fp32_peak, share_of_step=ki["pass1_ms"] / ms_step)
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=a.steps, warmup=a.warmup, ms_per_step=ms_step,
                higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32", data="synthetic",
                config=workload_config(a, world), e2e=e2e, gpu_launches=int(launches), clocks=clocks, roofline=roofline,
                stages=dict(pca_iterations=out["pca"]["iterations"], pca_converged=out["pca"]["converged"],
                            knn_uncertified_rows=ki["n_uncertified"], knn_resweep_rows=ki.get("n_resweep", 0), leiden=out["leiden_info"], n_communities=out["n_communities"],
                            modularity=out["modularity"]))
    if world == 1 and not a.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference_sample(a)
    emit(json.dumps(line))


_REAL_STDOUT = None


def emit(line: str) -> None:
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, library chatter) was diverted."""
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (line + "\n").encode())


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # native libraries that print to fd 1 (e.g. "NCCL version ...") now land on stderr
    a = parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if a.impl == "reference":
        run_reference(a, rank, world)
        return
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl")
    try:
        run_b200(a, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
