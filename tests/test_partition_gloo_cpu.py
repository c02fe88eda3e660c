"""Host-side logic of the multi-rank vector on CPU: world_size-2 gloo processes.

Each rank owns a contiguous block (sundials_b200.partition), computes the LOCAL
reductions with the CPU oracle, combines them with torch.distributed exactly as
the product does with NCCL (sum / max / min of 1..nv doubles; WRMS divides by
the GLOBAL length) and checks the result against the oracle on the whole vector.
No CUDA involved: this pins the partition + combine rules, not the kernels.
"""
import os
import subprocess
import sys
import textwrap
from pathlib import Path

import pytest

from sundials_b200.partition import COMBINE, block_range, local_length

ROOT = Path(__file__).resolve().parent.parent


def test_block_ranges_tile_the_vector():
    for n in (0, 1, 7, 8, 1000, 1 << 20):
        for size in (1, 2, 3, 8):
            stops = 0
            for r in range(size):
                a, b = block_range(n, r, size)
                assert a == stops and b >= a
                stops = b
                assert local_length(n, r, size) == b - a
            assert stops == n
            lens = [local_length(n, r, size) for r in range(size)]
            assert max(lens) - min(lens) <= 1
    with pytest.raises(ValueError):
        block_range(10, 2, 2)
    assert set(COMBINE.values()) == {"sum", "max", "min"}


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
    from _oracle import Oracle
    from sundials_b200.partition import block_range, COMBINE
    dist.init_process_group("gloo")
    rank, size = dist.get_rank(), dist.get_world_size()
    orc = Oracle()
    n = 100_003
    rng = np.random.default_rng(99)           # same global data on every rank
    x, y, w = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.uniform(0.5, 2, n)
    idv = rng.integers(0, 2, n).astype(np.float64)
    a, b = block_range(n, rank, size)
    xl, yl, wl, il = (np.ascontiguousarray(v[a:b]) for v in (x, y, w, idv))
    OPS = {{"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}}
    def combine(vals, how):
        t = torch.tensor(np.atleast_1d(np.asarray(vals, dtype=np.float64)))
        dist.all_reduce(t, op=OPS[how])
        return t.numpy()
    tol = 1e-13
    # sums: equal to the serial global result within the reduction tolerance
    g = combine(orc.dot_prod(xl, yl), COMBINE["dot_prod"])[0]
    assert abs(g - orc.dot_prod(x, y)) <= tol * np.abs(x * y).sum()
    g = combine(orc.wsqr_sum(xl, wl), COMBINE["wsqr_sum"])[0]
    wrms = 0.0 if g <= 0 else np.sqrt(g / n)             # GLOBAL length
    assert abs(wrms - orc.wrms_norm(x, w)) <= tol * wrms
    g = combine(orc.wsqr_sum_mask(xl, wl, il), COMBINE["wsqr_sum_mask"])[0]
    assert abs(np.sqrt(g / n) - orc.wrms_norm_mask(x, w, idv)) <= tol
    g = combine(orc.dot_prod_multi(xl, [yl, wl, xl]), COMBINE["dot_prod_multi"])
    assert np.all(np.abs(g - orc.dot_prod_multi(x, [y, w, x])) <= tol * n)
    # max / min / flags: exact
    assert combine(orc.max_norm(xl), COMBINE["max_norm"])[0] == orc.max_norm(x)
    assert combine(orc.min(xl), COMBINE["min"])[0] == orc.min(x)
    assert combine(orc.min_quotient(xl, yl), COMBINE["min_quotient"])[0] == orc.min_quotient(x, y)
    xz = x.copy(); xz[n - 1] = 0.0                      # a zero only on the last rank
    zl = np.empty(b - a); zg = np.empty(n)
    flag = combine(float(orc.inv_test(np.ascontiguousarray(xz[a:b]), zl)), COMBINE["inv_test"])[0]
    assert bool(flag) == orc.inv_test(xz, zg) == False
    # every rank holds identical global results (so integrators branch identically)
    t = torch.tensor([wrms]); lst = [torch.zeros(1, dtype=torch.float64) for _ in range(size)]
    dist.all_gather(lst, t.double())
    assert all(float(v) == float(lst[0]) for v in lst)
    dist.destroy_process_group()
    print("rank", rank, "OK")
""")


def test_world_size_2_gloo_combine_rules(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=str(ROOT)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("OK") == 2
