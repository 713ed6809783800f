"""Multi-GPU parity driver (one process per GPU; launched by torchrun from tests/test_gpu_multi.py or by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29541 tests/mgpu_parity.py [cells] [genes]

Every rank first runs the problem alone (no communicator: the single-GPU path), then the ranks create
the library communicator and run the SAME problem split across the GPUs.  Asserted on every rank:
identical iteration count, E and Z within 1e-10 relative, identical labels; the collective
eigensolver against the single-GPU one; the oracle on rank 0 when the size allows.
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rsoptsc_b200 import _lib, api, parallel  # noqa: E402
from rsoptsc_b200.synthetic import synthetic_lognorm  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 900
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 400
    rank, local_rank, world = parallel.world()
    lib = _lib.init(local_rank)
    dp = C.POINTER(C.c_double)
    data, _ = synthetic_lognorm(m, n, seed=3)
    X = api.normalize(data)
    pc = api.pca_spectrum(X, ncomp=3)
    idx = api.knn(pc["emb"], pc["K"])
    t0 = time.perf_counter()
    single = api.computeM(None, X, 0.5, idx=idx)
    t_single = time.perf_counter() - t0
    lab1 = api.nmf_lee(api.threshold_symmetrise(single["Z"]), 10)["labels"]
    rng = np.random.default_rng(5)
    ne = int(os.environ.get("MGPU_EIG_N", 700))
    B = rng.standard_normal((ne, ne // 2 + 3))
    G = np.asfortranarray(B @ B.T)
    lam1, V1 = np.zeros(ne), np.zeros((ne, ne), order="F")
    _lib.check(lib.rsoptsc_debug_sym_eig(G.ctypes.data_as(dp), ne, 1, lam1.ctypes.data_as(dp), V1.ctypes.data_as(dp), 1, None))

    parallel.init_comm()
    info0 = parallel.comm_info()
    assert info0["nranks"] == world and info0["rank"] == rank
    lam2, V2 = np.zeros(ne), np.zeros((ne, ne), order="F")
    ms = C.c_double(0)
    _lib.check(lib.rsoptsc_debug_sym_eig(G.ctypes.data_as(dp), ne, 2, lam2.ctypes.data_as(dp), V2.ctypes.data_as(dp), 1,
                                         C.byref(ms)))
    scale = np.abs(lam1).max()
    eig_err = np.abs(lam1 - lam2).max() / scale
    res = np.abs(G @ V2 - V2 * lam2).max() / scale
    orth = np.abs(V2.T @ V2 - np.eye(ne)).max()
    assert eig_err < 1e-12 and res < 1e-12 and orth < 1e-11, (eig_err, res, orth)

    t0 = time.perf_counter()
    multi = api.computeM(None, X, 0.5, idx=idx)
    t_multi = time.perf_counter() - t0
    lab2 = api.nmf_lee(api.threshold_symmetrise(multi["Z"]), 10)["labels"]
    info = parallel.comm_info()
    relZ = np.linalg.norm(multi["Z"] - single["Z"]) / np.linalg.norm(single["Z"])
    relE = abs(multi["E"] - single["E"]) / single["E"]
    assert multi["iters"] == single["iters"], (multi["iters"], single["iters"])
    assert relZ < 1e-10 and relE < 1e-12, (relZ, relE)
    assert np.array_equal(lab1, lab2)
    assert world == 1 or info["collectives"] > info0["collectives"]
    out = dict(rank=rank, world=world, cells=n, genes=m, K=int(pc["K"]), iters=multi["iters"], relZ=relZ, relE=relE,
               eig_err=eig_err, eig_res=res, eig_orth=orth, eig_ms=ms.value, p2p=info["p2p"],
               collectives=info["collectives"], collective_GB=info["bytes"] / 1e9, t_single_s=t_single, t_multi_s=t_multi)
    if rank == 0 and n <= 1500:
        from oracle import rsoptsc_oracle as O
        ref = O.computeM(O.mask_from_idx(idx, n), X, 0.5, literal=False)
        out["relZ_oracle"] = float(np.linalg.norm(multi["Z"] - ref["Z"]) / np.linalg.norm(ref["Z"]))
        assert ref["iters"] == multi["iters"] and out["relZ_oracle"] < 1e-6
    print("MGPU_PARITY " + json.dumps(out), flush=True)
    parallel.finalize_comm()


if __name__ == "__main__":
    main()
