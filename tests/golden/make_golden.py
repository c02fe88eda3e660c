"""Generate tests/golden/nvec_golden.npz from the UNMODIFIED reference.

Runs every parity case of tests/_cases.py through the reference's own
nvector_serial (oracle/_ref/lib/libsundials_ref.so, compiled from
/root/reference by oracle/Makefile) and stores the outputs.  The inputs are
re-created from the seeds below by the tests, so only outputs are stored.

    make -C oracle ref && python tests/golden/make_golden.py

This can only run where /root/reference (or a prebuilt oracle/_ref) exists; the
resulting .npz is committed and is what pins oracle/nvec_oracle.c on machines
without the reference.
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from _cases import all_cases  # noqa: E402
from _oracle import RefSerial  # noqa: E402

GOLDEN_N = (67, 1)
GOLDEN_SEED = 20240607


def main():
    ref = RefSerial()
    out = {}
    for n in GOLDEN_N:
        for name, fn in all_cases():
            if n == 1 and not name.startswith(("linear_sum[", "scale[", "dot_prod", "min", "linear_combination[")):
                continue
            res = fn(ref, n, GOLDEN_SEED)
            for k, v in res.items():
                out[f"n{n}|{name}|{k}"] = np.atleast_1d(np.asarray(v, dtype=np.float64))
    path = Path(__file__).resolve().parent / "nvec_golden.npz"
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {path.stat().st_size/1024:.0f} KiB")


if __name__ == "__main__":
    main()
