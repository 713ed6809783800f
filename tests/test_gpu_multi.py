"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): launches tests/mgpu_parity.py under
torchrun at world size 2 and checks what every rank reports -- the sharded computeM (column blocks of
the ADMM state, collective SVT, NCCL allreduce of the stop-rule Gram) equals the single-GPU run."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=20).stdout
        return sum(1 for line in out.splitlines() if line.startswith("GPU "))
    except Exception:
        return 0


def _run(world, cells, genes, env_extra, port):
    env = dict(os.environ, **env_extra)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_parity.py"),
           str(cells), str(genes)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    rows = [json.loads(line.split("MGPU_PARITY ", 1)[1]) for line in r.stdout.splitlines() if "MGPU_PARITY " in line]
    assert len(rows) == world
    return rows


@pytest.mark.gpu
def test_sharded_computeM_equals_single_gpu_world2():
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    # small thresholds so that the 900-cell problem exercises every split path
    # MGPU_EIG_N = 1500: the collective eigensolver test crosses several 128-row ownership groups, panels that start
    # on odd and even multiples of 64 (lead rows of the sharded trailing update) and the small-block tail
    rows = _run(2, 900, 400, dict(RSOPTSC_SPLIT_MIN_N="150", RSOPTSC_STEDC_SPLIT_MIN="128", RSOPTSC_SYTRD_MG_MIN_N="128",
                                  RSOPTSC_GEMM_MIN="0", MGPU_EIG_N="1500"), 29541)
    for r in rows:
        assert r["relZ"] < 1e-10 and r["collectives"] > 0
    assert "relZ_oracle" in [k for r in rows for k in r]
