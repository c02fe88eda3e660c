"""bench.py's output contract on a GPU: one JSON line with the keys the driver parses, the roofline /
cpu_baseline / e2e objects, and the compact `legs` record LAST (it must survive a tail cut of the line)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_bench_line_contract_small_run():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--log2n", "20", "--cpu-log2n", "16", "--steps", "3",
                        "--warmup", "3", "--no-diffusion", "--no-ar3d", "--no-gs"],
                       capture_output=True, text=True, timeout=900, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks", "legs"):
        assert k in d, k
    assert list(d)[-1] == "legs"
    assert d["dtype"] == "f64" and d["unit"] == "GB/s" and d["n_gpus"] == 1 and d["gpu_launches"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 8 * (1 << 20) * 8 and d["e2e"]["d2h_bytes_per_step"] > 0
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-3
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] > 0
    # the in-bench parity proof against the reference's own CPU vector, through the same C driver
    assert cb["checksum_parity"]["ok"] is True, cb["checksum_parity"]
    assert d["legs"]["checksum_vs_reference_ok"] is True
    assert d["cvDiurnal_kry"]["stdout_identical_to_serial_golden"] is True
    assert set(d["sweep"]["lengths"]) == {"2^16", "2^20", "2^24", "2^28"}
    assert len(json.dumps(d["legs"])) < 1400
    assert len(d["legs"]["cvode_fused"]) == 7 and all(f > 0 for f, _ in d["legs"]["cvode_fused"].values())
