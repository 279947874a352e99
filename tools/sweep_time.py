"""Step + projection timing of selected sweep cases (bench.projection_sweep), for kernel experiments:
    [PAA_LIBPAA=/path/to/variant.so] python tools/sweep_time.py [case ...]      (no case = all, "compose" = compose kernels)
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

only = set(sys.argv[1:]) or None
res = bench.projection_sweep(torch.device("cuda:0"), iters=30, only=only)
tag = os.path.basename(os.environ.get("PAA_LIBPAA", "libpaa.so"))
for k, v in res.items():
    print(tag, k, json.dumps({q: v[q] for q in ("shape", "p_rows", "ms", "wall_ms", "GB/s", "frac_of_measured_peak", "gflops", "frac_of_fp32_peak") if q in v}))
