#!/usr/bin/env python
"""Benchmark of the RSoptSC hot path: SimilarityM + ClusterCells (k = 10).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of SimilarityM -> ClusterCells over one synthetic log-normalised
genes x cells matrix (BASELINE.json configs[1]: 10,000 cells x 2,000 genes on 1 x B200).
  value  whole-job cells/s with the input matrix resident in HBM (rsoptsc_*_dev entry points)
  e2e    the same through the host-buffer C ABI (rsoptsc_similarity + rsoptsc_nmf_lee): pinned
         host input, host<->device copies inside the timed region
  N > 1  (torchrun) ONE problem split across the N GPUs inside the library (librsoptsc_b200's own NCCL
         communicator: column blocks of the ADMM state, collective SVT with the Gram / eigensolver /
         reconstruction split across the ranks, NCCL allreduce of the stop-rule Grams) -> strong scaling;
         --replicas keeps the round-1 mode (N independent matrices, weak scaling)
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
CPU_SAMPLE_CELLS = 1000   # BASELINE.json configs[0]: the reference's own CPU-runnable case, run in full
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


def best_cpu_threads(data, cells=500):
    """BLAS thread count for the CPU arm.  torchrun pins OMP_NUM_THREADS=1 and an oversubscribed OpenBLAS can be slower
    than a few threads, so a short probe at 500 cells (large enough for the threaded LAPACK paths to matter: a
    300-cell probe picked 1 thread on a 16-CPU box where 8 threads are 1.9x faster at 1,000 cells) times the multi-thread
    candidates {all, half, 8} host CPUs; a single thread is only a candidate on a single-CPU host."""
    ncpu = os.cpu_count() or 1
    cand = sorted({min(8, ncpu), max(1, ncpu // 2), ncpu}, reverse=True)
    if ncpu >= 2:
        cand = [c for c in cand if c >= 2]
    cells = min(cells, data.shape[1])
    cpu_oracle_sample(data, min(cells, 200), steps=1, threads=cand[0])      # page-in / thread-pool warm-up
    best = None
    for th in cand:
        t = min(cpu_oracle_sample(data, cells, steps=1, threads=th))
        if best is None or t < best[1]:
            best = (th, t)
    return best


def bench_config(args, world):
    """The workload both arms describe (identical dict in the GPU line and in the reference line)."""
    if args.replicas and world > 1:
        par = "replicas x%d (no data-path collective)" % world
    elif world > 1:
        par = ("1 problem over %d GPUs: cells (columns of Z, E, Y1, XXZ, Y3, J) sharded, SVT collective (Gram panels "
               "+ eigensolver stages + reconstruction split, NCCL all-gather / allreduce over NVLink)" % world)
    else:
        par = "1 GPU"
    return {"workload": workload_name(args), "cells": args.cells, "genes": args.genes, "lambda": LAMBDA,
            "nmf_rank": RANK, "parallelism": par,
            "headline_config_measured": args.cells >= 50000,
            "headline_note": ("BASELINE.json's metric is quoted at 50,000 cells (configs[2]); one step of it takes 443 s "
                              "on one B200, 296 s on 2, 161 s on 4 and 102 s on 8 (profiles/r2_bench_50k_*gpu.json, this same code with "
                              "`--cells 50000 --steps 1 --warmup 0`), so the driver-sized run (--steps 20) uses configs[1]"),
            "l2": "working set (%.1f GB of dense n x n state) exceeds the 126 MB L2" % (5 * 8e-9 * args.cells * args.cells)}


def run_reference(args, rank, world):
    """Reference arm: the CPU oracle (numpy/OpenBLAS restatement of R/SimilarityM.R + R/Clustering.R; R itself
    is not installable here), all useful host threads.  Every step is the SAME pipeline as the GPU arm's step on a
    bounded sample of the workload: the first `cells` columns of the seed-0 matrix.  cells = 1,000 (BASELINE.json
    configs[0] in full) when steps x its time fits ~6 minutes, else 700 / 500 -- stated in cpu_baseline.sample."""
    if rank != 0:
        return
    from rsoptsc_b200.synthetic import synthetic_lognorm
    data, _ = synthetic_lognorm(args.genes, 2000, seed=0)
    threads, t500 = best_cpu_threads(data, min(500, args.cpu_cells) if args.cpu_cells_fixed else 500)  # + BLAS warm-up
    cells = args.cpu_cells
    if not args.cpu_cells_fixed:
        for cand in (1000, 700, 500):
            cells = cand
            if args.steps * t500 * (cand / 500.0) ** 2.6 <= 360.0:
                break
    for _ in range(min(args.warmup, 1)):            # warm-up steps: one short pass (page-in, thread pool)
        cpu_oracle_sample(data, min(cells, 300), threads=threads)
    times = cpu_oracle_sample(data, cells, steps=args.steps, threads=threads)
    t = float(np.mean(times))
    value = cells / t
    sample = ("each step = full SimilarityM (literal full-SVD norms as in R) + ClusterCells k=%d on the first %d cells x "
              "%d genes of the seed-0 synthetic matrix (%s), %d BLAS threads of %d host CPUs; cost grows ~n^2.6-n^3, so "
              "cells/s at %d cells is far lower (not measured: hours)"
              % (RANK, cells, args.genes, "BASELINE.json configs[0] in full" if cells == 1000 else "bounded sample",
                 threads, os.cpu_count() or 1, args.cells))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak" if args.replicas else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count(), "sample_cells": cells,
                         "reference_arm": "numpy/OpenBLAS oracle (R is not installable here)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_name(args):
    cfg = {1000: "configs[0]", 10000: "configs[1]", 50000: "configs[2]"}.get(args.cells, "custom")
    return ("%s: %d cells x %d genes synthetic log-normalised, SimilarityM (lambda=%.2g, in-library PCA "
            "embedding) + ClusterCells (NMF lee/nndsvd, k=%d)" % (cfg, args.cells, args.genes, LAMBDA, RANK))


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
    ap.add_argument("--cpu-cells", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--replicas", action="store_true", help="N > 1: independent matrices per GPU (weak scaling)")
    args = ap.parse_args()
    args.cpu_cells_fixed = args.cpu_cells is not None
    if args.cpu_cells is None:
        args.cpu_cells = CPU_SAMPLE_CELLS
    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from rsoptsc_b200 import _lib, api, parallel
    from rsoptsc_b200.synthetic import synthetic_lognorm

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the rsoptsc_b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.init(local_rank)
    shard = world > 1 and not args.replicas
    if shard:
        # the library's own NCCL communicator (no torch inside the package): rank 0's unique id travels over
        # the process group torchrun already gave this script
        def bcast(payload):
            box = [payload]
            dist.broadcast_object_list(box, src=0)
            return box[0]
        parallel.init_comm(rank, world, device=local_rank, broadcast=bcast)
    m, n = args.genes, args.cells

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # strong scaling: the SAME matrix on every rank (inputs are replicated, the state is sharded);
    # --replicas: one independent matrix per rank
    data, _ = synthetic_lognorm(m, n, seed=0 if not args.replicas else rank)
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
    comm0 = parallel.comm_info() if shard else None
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
    comm1 = parallel.comm_info() if shard else None
    oz_flops = lib.rsoptsc_counter(b"ozaki_fp64_flops")
    oz_iops = lib.rsoptsc_counter(b"ozaki_int8_ops")
    sytrd_bytes = lib.rsoptsc_counter(b"sytrd_panel_bytes")
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

    # the same pipeline at the size the CPU oracle is timed on (configs[0], 1,000 cells): the one like-for-like
    # GPU / CPU pair of this report (rank 0, N = 1 only; untimed part of the run)
    same_size = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.cpu_cells < n:
        nc = args.cpu_cells
        sub = np.asfortranarray(data[:, :nc])
        d_sub = api.DeviceBuffer.from_host(sub)
        d_Ws = api.DeviceBuffer(nc * nc * 8)

        def step_small():
            api.SimilarityM_dev(LAMBDA, d_sub, m, nc, d_W=d_Ws)
            api.nmf_lee_dev(d_Ws, nc, RANK)
        step_small()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            step_small()
        lib.rsoptsc_dev_sync()
        same_size = {"cells": nc, "gpu_s_per_step": (time.perf_counter() - t0) / 3}

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
        bf16 = peaks.get("bf16_tflops_sustained") or 1400.0
        bf16_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks.get("bf16_tflops_sustained") else \
            "fallback (B200_PROFILING.md)"
        nnz = n * info.get("K", 30)
        roof = None
        oz = phases.get("ozaki_gemm")
        sy = phases.get("sytrd_panel")

        def ozaki_roof():
            pairs = max(oz_iops / max(oz_flops, 1.0), 1.0)
            osec = oz["ms"] / 1e3
            return {"kernel": "k_ozaki_gemm<8> (tcgen05.mma kind::i8, FP64-equivalent)", "bound": "tensor",
                    "achieved": oz_flops / osec / 1e12, "peak": 2.0 * bf16 / pairs, "unit": "TFLOP/s",
                    "frac": (oz_flops / osec / 1e12) / (2.0 * bf16 / pairs),
                    "traffic": 15.35e9,
                    "traffic_note": "dram read+write of ONE launch, the n=10000 Gram, from ncu --set full "
                                    "(profiles/r2_ozaki_gemm_ncu_full.csv): 14.5 GB read + 0.8 GB written, DRAM at 14 %, "
                                    "tensor pipe active 72 %, bound by L2 -> SM operand traffic",
                    "int8_tops_achieved": oz_iops / osec / 1e12, "int8_tops_peak": 2.0 * bf16,
                    "peak_source": "2 x %s (int8 runs at twice the bf16 rate) / %d int8 slice-pair MMAs per FP64 MAC"
                                   % (bf16_src, round(pairs)),
                    "share_of_step": osec / t_dev, "launches": oz["calls"], "cublas_fp64_reference_tflops": 35.6}

        if sy and sy["ms"] > 0 and (not oz or sy["ms"] >= oz["ms"]):
            # dominant kernel: the cooperative panel kernel of the tridiagonalisation (symmetric eigensolver of
            # every SVT).  HBM-bound: its product phase reads the lower triangle of the trailing matrix once per
            # column; algorithmic bytes = sum over columns of 4 (n - j - 1)^2 (this rank's share when sharded).
            secs = sy["ms"] / 1e3
            achieved = sytrd_bytes / secs / 1e9
            roof = {"kernel": "k_latrd_panel%s (Householder tridiagonalisation, 64-column panels)" % ("_mg" if shard else ""),
                    "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": SYTRD_NCU_TRAFFIC if not shard else None,
                    "traffic_note": SYTRD_NCU_NOTE if not shard else "rank 0's share of the lower triangles",
                    "peak_source": peak_src, "share_of_step": secs / t_dev, "launches": sy["calls"],
                    "algorithmic_bytes_per_launch": sytrd_bytes / max(sy["calls"], 1)}
            if oz and oz["ms"] > 0:
                roof["second_kernel"] = ozaki_roof()
        elif oz and oz["ms"] > 0:
            roof = ozaki_roof()
        else:
            own = [(p, v) for p, v in phases.items() if own_kernel_bytes(p, m, n, nnz)]
            if own:
                name, v = max(own, key=lambda kv: kv[1]["ms"])
                per_launch_s = v["ms"] / 1e3 / max(v["calls"], 1)
                achieved = own_kernel_bytes(name, m, n, nnz) / per_launch_s / 1e9
                roof = {"kernel": name, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                        "frac": achieved / hbm_peak, "traffic": None, "peak_source": peak_src,
                        "share_of_step": v["ms"] / 1e3 / t_dev, "launches": v["calls"]}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            sdata = data if n >= 300 else synthetic_lognorm(m, 300, seed=0)[0]
            cpu_threads, _ = best_cpu_threads(sdata, min(500, n))
            nc = min(args.cpu_cells, n)
            t_cpu = cpu_oracle_sample(data, nc, threads=cpu_threads)[0]
            cpu = {"value": nc / t_cpu, "unit": UNIT, "cores": cpu_threads, "host_cpus": os.cpu_count(), "kind": "port",
                   "sample": "first %d cells x %d genes of the same matrix (%s), oracle SimilarityM (literal full-SVD "
                             "norms as in R) + ClusterCells k=%d once: %.1f s; cost ~n^3, so this overstates CPU cells/s "
                             "at %d cells" % (nc, m, "BASELINE.json configs[0] in full" if nc == 1000 else "bounded sample",
                                              RANK, t_cpu, n),
                   "seconds": t_cpu, "sample_cells": nc,
                   "extrapolated_seconds_at_bench_size": t_cpu * (n / nc) ** 3,
                   "extrapolation": "n^3 law from the one measured point; EXTRAPOLATED, not measured"}
            if same_size:
                same_size.update(cpu_s_per_step=t_cpu, gpu_cells_per_s=nc / same_size["gpu_s_per_step"],
                                 cpu_cells_per_s=nc / t_cpu, gpu_over_cpu=t_cpu / same_size["gpu_s_per_step"],
                                 note="same pipeline, same %d-cell input, same box: the like-for-like GPU/CPU pair" % nc)
        value = (world if args.replicas else 1) * n * args.steps / t_dev
        cfg = bench_config(args, world)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak" if args.replicas else "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "wall_s_per_step": t_wall / args.steps,
            "e2e": ({"value": (world if args.replicas else 1) * n * args.steps / t_e2e, "unit": UNIT,
                     "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                     "ms_per_step": t_e2e / args.steps * 1e3} if e2e else None),
            "gpu_launches": int(launches),
            "roofline": roof,
            "cpu_baseline": cpu,
            "clocks": sampler.summary(),
            "extra": {"K": info.get("K"), "admm_iters": info.get("iters"), "nmf_iters": info.get("nmf_iters"),
                      "E": info.get("E"), "same_size_as_cpu_baseline": same_size,
                      "comm": ({"nranks": comm1["nranks"], "peer_mapped_exchange": comm1["p2p"],
                                "collectives_per_step": (comm1["collectives"] - comm0["collectives"]) / args.steps,
                                "GB_per_step_rank0": (comm1["bytes"] - comm0["bytes"]) / 1e9 / args.steps}
                               if shard else None)},
            "phases_ms_per_step": {p: round(v["ms"] / args.steps, 3) for p, v in phases.items() if p not in ("e2e",)},
        }
        print(json.dumps(out))
    if shard:
        parallel.finalize_comm()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
