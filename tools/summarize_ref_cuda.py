"""profiles/*_ref_cuda_suite.json -> markdown table.  python tools/summarize_ref_cuda.py in.json > out.md"""
import json
import sys

d = json.load(open(sys.argv[1]))
print(f"# NVECTOR_B200 vs the reference's nvector_cuda recompiled for sm_100a — same {d['gpu']}, same process\n")
print("`tools/ref_cuda_suite.py`: identical 55-op suite (benchmarks/nvector cases), n = 2^%d, fused ops enabled on both,"
      " same seeded inputs; timing = %s (%d reps/op).\n" % (d["log2n"], d["timing"], d["reps"]))
print(f"Whole suite: nvector_cuda {d['suite_ms']['ref_cuda']} ms ({d['suite_GBs']['ref_cuda']} GB/s algorithmic) vs "
      f"NVECTOR_B200 {d['suite_ms']['b200']} ms ({d['suite_GBs']['b200']} GB/s): **{d['suite_speedup']}x**.\n")
print("| op | B/elt | nvector_cuda us | B200 us | nvector_cuda GB/s | B200 GB/s | speed-up |")
print("|---|---:|---:|---:|---:|---:|---:|")
for k, v in d["per_op"].items():
    print(f"| {k} | {v['B_per_elt']} | {v['ref_cuda_us']} | {v['b200_us']} | {v['ref_cuda_GBs']} | {v['b200_GBs']} | {v['speedup']} |")
print("\nScalar results of the two arms on the same inputs (reference, ours):\n")
for k, (r, b) in d["scalars_ref_vs_b200"].items():
    rel = abs(r - b) / max(abs(r), 1e-300)
    print(f"* {k}: {r!r} vs {b!r} (rel. diff {rel:.1e})")
