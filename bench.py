#!/usr/bin/env python
"""Benchmark of the RSoptSC hot path: SimilarityM + ClusterCells (k = 10).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of SimilarityM -> ClusterCells over one synthetic log-normalised
genes x cells matrix (BASELINE.json configs[1]: 10,000 cells x 2,000 genes on 1 x B200).
  value  whole-job cells/s with the input matrix resident in HBM (rsoptsc_*_dev entry points)
  e2e    the same through the host-buffer C ABI (rsoptsc_similarity + rsoptsc_nmf_lee): pinned
         host input, host<->device copies inside the timed region
  N > 1  (torchrun) N independent replicas, one matrix per GPU, no data-path collective
         ("replicas only" this round) -> weak scaling
  --impl reference   the CPU oracle (numpy/OpenBLAS restatement of the R code, all host
         threads) on a bounded sample of the same workload; rank 0 only
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_GENES = 2000
N_CELLS = 10000
LAMBDA = 0.5
RANK = 10
CPU_SAMPLE_CELLS = 500
METRIC = "SimilarityM+ClusterCells throughput"
UNIT = "cells/s"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    f = [x.strip() for x in line.split(",")]
                    if len(f) >= 8:
                        self.rows.append(f)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(r[4 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def cpu_oracle_sample(data, cells, steps=1, threads=None):
    """Time the CPU oracle (reference restatement) on the first `cells` columns with `threads`
    BLAS threads (None: library default)."""
    from oracle import rsoptsc_oracle as O
    sub = np.asfortranarray(data[:, :cells])
    times = []

    def run():
        for _ in range(steps):
            t0 = time.perf_counter()
            r = O.SimilarityM(LAMBDA, sub, literal=True)
            O.ClusterCells(r["W"], n_clusters=RANK)
            times.append(time.perf_counter() - t0)
    if threads:
        from threadpoolctl import threadpool_limits
        with threadpool_limits(limits=int(threads)):
            run()
    else:
        run()
    return times


def best_cpu_threads(data, cells):
    """torchrun pins OMP_NUM_THREADS=1 and an oversubscribed OpenBLAS can be slower than one thread
    at this size: time the sample with all host threads and with one, keep the faster setting."""
    ncpu = os.cpu_count() or 1
    cand = sorted({1, ncpu})
    best = None
    for th in cand:
        t = cpu_oracle_sample(data, cells, threads=th)[0]
        if best is None or t < best[1]:
            best = (th, t)
    return best


def run_reference(args, rank):
    if rank != 0:
        return
    from rsoptsc_b200.synthetic import synthetic_lognorm
    cells = args.cpu_cells
    data, _ = synthetic_lognorm(args.genes, max(cells, 16), seed=0)
    threads, _ = best_cpu_threads(data, cells)  # doubles as the warm-up pass; picks 1 vs all host threads
    times = cpu_oracle_sample(data, cells, steps=args.steps, threads=threads)
    t = float(np.mean(times))
    value = cells / t
    sample = ("first %d cells x %d genes of the seed-0 synthetic matrix, full SimilarityM (literal full-SVD norms as "
              "in R) + ClusterCells k=%d; cost grows ~n^3 so cells/s at %d cells would be far lower"
              % (cells, args.genes, RANK, args.cells))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "cells": args.cells, "genes": args.genes, "lambda": LAMBDA,
                   "nmf_rank": RANK, "reference_arm": "numpy/OpenBLAS oracle (R is not installable here)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_name(args):
    return ("configs[1]: %d cells x %d genes synthetic log-normalised, SimilarityM (lambda=%.2g, in-library PCA "
            "embedding) + ClusterCells (NMF lee/nndsvd, k=%d) on 1 B200 per replica"
            % (args.cells, args.genes, LAMBDA, RANK))


# algorithmic bytes per launch of the HBM-bound kernels we own
def own_kernel_bytes(name, m, n, nnz):
    return {
        "update_Y3": 24.0 * n * n,                 # read J, read+write Y3
        "build_A": 16.0 * n * n + 8.0 * n * n,     # read Y3, write A, re-read A for ||A||_F
        "E_update": 32.0 * m * n,                  # read XXZ, Y1; write E, R
        "Z_update": 8.0 * m * nnz + 8.0 * m * n,   # one X column per support entry + R once
        "residual": 8.0 * m * nnz + 48.0 * m * n,  # X columns + X, E, Y1 in; XXZ, Y1, M out
        "symm": 16.0 * n * n,                      # memset + scatter of the dense W
    }.get(name)


# ncu --set full of one k_latrd_panel launch (first panel of a 7,954-row block): filled from profiles/
SYTRD_NCU_TRAFFIC = 17.32e9
SYTRD_NCU_NOTE = ("dram read+write of ONE launch, the first 64-column panel of a 7,954-row block, from ncu --set full "
                  "(profiles/r1_latrd_panel_ncu_full.csv): 17.14 GB read + 0.18 GB written against 16.06 GB algorithmic "
                  "(lower triangles of the 64 trailing matrices) = 1.08x; 4.78 TB/s DRAM read rate over the launch; the "
                  "rest of the time is the two grid synchronisations and the row phase of every column")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native")
    ap.add_argument("--cells", type=int, default=N_CELLS)
    ap.add_argument("--genes", type=int, default=M_GENES)
    ap.add_argument("--cpu-cells", type=int, default=CPU_SAMPLE_CELLS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from rsoptsc_b200 import _lib, api
    from rsoptsc_b200.synthetic import synthetic_lognorm

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the rsoptsc_b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    _lib.init(local_rank)
    m, n = args.genes, args.cells

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # one independent matrix per replica (seed = rank)
    data, _ = synthetic_lognorm(m, n, seed=rank)
    d_data = api.DeviceBuffer.from_host(data)
    d_W = api.DeviceBuffer(n * n * 8)
    info = {}

    def step_resident():
        r = api.SimilarityM_dev(LAMBDA, d_data, m, n, d_W=d_W)
        c = api.nmf_lee_dev(d_W, n, RANK)
        info.update(iters=r["iters"], K=r["K"], nmf_iters=c["iters"], E=r["E"])
        return c["labels"]

    for _ in range(args.warmup):
        step_resident()
    api.timers(True)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:  # one nvidia-smi poller per job is enough (rank 0's GPU)
        sampler.start()
    launches0 = api.kernel_launches()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        api.timer_begin("step")
        step_resident()
        api.timer_end("step")
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = api.kernel_launches() - launches0
    t_dev = api.phase_ms("step") / 1e3
    phases = api.phase_report()
    oz_flops = _lib.load().rsoptsc_counter(b"ozaki_fp64_flops")
    oz_iops = _lib.load().rsoptsc_counter(b"ozaki_int8_ops")
    sytrd_bytes = _lib.load().rsoptsc_counter(b"sytrd_panel_bytes")
    api.timers(False)

    # end to end through the host-buffer C ABI, pinned host input/output
    e2e = None
    if not args.no_e2e:
        h_data = torch.from_numpy(data.T.copy()).pin_memory()      # (n, m) C-order == (m, n) Fortran-order
        h_W = torch.empty((n, n), dtype=torch.float64).pin_memory()
        data_pinned = h_data.numpy().T
        W_pinned = h_W.numpy().T

        def step_e2e():
            r = api.SimilarityM(LAMBDA, data_pinned, W_out=W_pinned)
            c = api.nmf_lee(r["W"], RANK)
            return c["labels"]

        step_e2e()
        barrier()
        t0 = time.perf_counter()
        api.timer_begin("e2e")
        for _ in range(args.steps):
            step_e2e()
        api.timer_end("e2e")
        barrier()
        t_e2e_wall = time.perf_counter() - t0
        t_e2e = max(api.phase_ms("e2e") / 1e3, 0.0)
        e2e = dict(t=t_e2e, wall=t_e2e_wall,
                   h2d=8.0 * m * n + 8.0 * n * n, d2h=8.0 * n * n + 8.0 * n * RANK * 2 + 4.0 * n)
    sampler.stop_flag.set()
    if sampler.is_alive():
        sampler.join(timeout=2)

    # max over ranks
    tt = torch.tensor([t_dev, e2e["t"] if e2e else 0.0, t_wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_dev, t_e2e, t_wall = [float(x) for x in tt.tolist()]

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak, peak_src = (peaks.get("hbm_gbs"), "MEASURED_PEAKS.json hbm_gbs") if peaks.get("hbm_gbs") else \
            (6650.0, "fallback (B200_PROFILING.md)")
        nnz = n * info.get("K", 30)
        # dominant kernel among the kernels this repo owns (cuSOLVER/cuBLAS phases reported in `phases`)
        own = [(p, v) for p, v in phases.items() if own_kernel_bytes(p, m, n, nnz)]
        roof = None
        oz = phases.get("ozaki_gemm")
        sy = phases.get("sytrd_panel")
        if sy and sy["ms"] > 0 and (not oz or sy["ms"] >= oz["ms"]):
            # dominant kernel: the cooperative panel kernel of the tridiagonalisation (symmetric
            # eigensolver of every SVT).  HBM-bound: its product phase reads the lower triangle of the
            # trailing matrix once per column; algorithmic bytes = sum over columns of 4 (n - j - 1)^2.
            secs = sy["ms"] / 1e3
            achieved = sytrd_bytes / secs / 1e9
            roof = {"kernel": "k_latrd_panel (Householder tridiagonalisation, 64-column panels)", "bound": "hbm",
                    "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": SYTRD_NCU_TRAFFIC,
                    "traffic_note": SYTRD_NCU_NOTE,
                    "peak_source": peak_src, "share_of_step": secs / t_dev, "launches": sy["calls"],
                    "algorithmic_bytes_per_launch": sytrd_bytes / max(sy["calls"], 1)}
            if oz and oz["ms"] > 0:
                bf16 = peaks.get("bf16_tflops_sustained") or 1400.0
                pairs = max(oz_iops / max(oz_flops, 1.0), 1.0)
                osec = oz["ms"] / 1e3
                roof["second_kernel"] = {
                    "kernel": "k_ozaki_gemm<8> (tcgen05.mma kind::i8, FP64-equivalent)", "bound": "tensor",
                    "achieved": oz_flops / osec / 1e12, "peak": 2.0 * bf16 / pairs, "unit": "TFLOP/s",
                    "frac": (oz_flops / osec / 1e12) / (2.0 * bf16 / pairs), "share_of_step": osec / t_dev,
                    "launches": oz["calls"], "cublas_fp64_reference_tflops": 35.6}
        elif oz and oz["ms"] > 0:
            # dominant kernel this repo owns: the tcgen05 kind::i8 split-integer GEMM (Gram + the two
            # reconstruction GEMMs of every SVT).  achieved = FP64-equivalent flops / device time;
            # peak = int8 tensor peak (2 x measured dense bf16) / 36 slice-pair MMAs per FP64 MAC.
            bf16 = peaks.get("bf16_tflops_sustained") or 1400.0
            bf16_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks.get("bf16_tflops_sustained") else \
                "fallback (B200_PROFILING.md)"
            flops, iops = oz_flops, oz_iops
            secs = oz["ms"] / 1e3
            pairs = max(iops / max(flops, 1.0), 1.0)
            roof = {"kernel": "k_ozaki_gemm<8> (tcgen05.mma kind::i8, FP64-equivalent)", "bound": "tensor",
                    "achieved": flops / secs / 1e12, "peak": 2.0 * bf16 / pairs, "unit": "TFLOP/s",
                    "frac": (flops / secs / 1e12) / (2.0 * bf16 / pairs),
                    "traffic": 13.43e9,
                    "traffic_note": "dram read+write of ONE launch, the n=10000 Gram, from ncu --set full "
                                    "(profiles/r1_ozaki_gemm_ncu_full.csv); its operands are 0.8 GB of int8 slices + "
                                    "0.4 GB FP64 output, re-read per tile row/column through L2 (11 % of DRAM "
                                    "bandwidth: the kernel is tensor-bound, tensor pipe active 65 %)",
                    "int8_tops_achieved": iops / secs / 1e12, "int8_tops_peak": 2.0 * bf16,
                    "frac_of_nominal_int8_4500_tops": iops / secs / 1e12 / 4500.0,
                    "peak_source": "2 x %s (int8 runs at twice the bf16 rate) / %d int8 slice-pair MMAs per FP64 MAC"
                                   % (bf16_src, round(pairs)),
                    "share_of_step": secs / t_dev, "launches": oz["calls"],
                    "cublas_fp64_reference_tflops": 35.6}
        elif own:
            name, v = max(own, key=lambda kv: kv[1]["ms"])
            per_launch_s = v["ms"] / 1e3 / max(v["calls"], 1)
            achieved = own_kernel_bytes(name, m, n, nnz) / per_launch_s / 1e9
            roof = {"kernel": name, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                    "frac": achieved / hbm_peak, "traffic": None, "peak_source": peak_src,
                    "share_of_step": v["ms"] / 1e3 / t_dev, "launches": v["calls"]}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cpu_threads, t_cpu = best_cpu_threads(data, args.cpu_cells)
            cpu = {"value": args.cpu_cells / t_cpu, "unit": UNIT, "cores": cpu_threads, "host_cpus": os.cpu_count(),
                   "kind": "port",
                   "sample": "first %d cells x %d genes of the same matrix, oracle SimilarityM (literal full-SVD norms) "
                             "+ ClusterCells k=%d, %.1f s; cost ~n^3, so this overstates CPU cells/s at %d cells"
                             % (args.cpu_cells, m, RANK, t_cpu, n), "seconds": t_cpu}
        value = world * n * args.steps / t_dev
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "cells": n, "genes": m, "lambda": LAMBDA, "nmf_rank": RANK,
                       "K": info.get("K"), "admm_iters": info.get("iters"), "nmf_iters": info.get("nmf_iters"),
                       "parallelism": "replicas x%d (no data-path collective)" % world,
                       "l2": "working set (%.1f GB dense state per replica) exceeds the 126 MB L2" % (5 * 8e-9 * n * n)},
            "wall_s_per_step": t_wall / args.steps,
            "e2e": ({"value": world * n * args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"],
                     "d2h_bytes_per_step": e2e["d2h"], "ms_per_step": t_e2e / args.steps * 1e3} if e2e else None),
            "gpu_launches": int(launches),
            "roofline": roof,
            "cpu_baseline": cpu,
            "clocks": sampler.summary(),
            "phases_ms_per_step": {p: round(v["ms"] / args.steps, 3) for p, v in phases.items() if p not in ("e2e",)},
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
