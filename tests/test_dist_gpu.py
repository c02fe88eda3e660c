"""Distributed NVECTOR_B200 on N GPUs of one box (SURVEY.md rows a20 / e): runs
tests/dist_parity_gpu.py under torchrun -- every reducing op of a vector partitioned in
contiguous blocks (MPIPlusX pattern) against the CPU oracle on the WHOLE vector, both
transports (in-kernel fold over NVLink peer memory, ncclAllReduce), empty local blocks,
identical scalars on every rank.  Skipped when the box has fewer GPUs than ranks.
"""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_distributed_reductions_match_oracle_on_global_vector(nproc):
    import torch

    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29544 + nproc),
                        str(ROOT / "tests" / "dist_parity_gpu.py")],
                       capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert f"DIST PARITY OK world={nproc}" in r.stdout
