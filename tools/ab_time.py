"""Fine-grained A/B timer for kernel experiments (CUDA events on this box tick every 2.048 us, too coarse for one call).

Each case is run back to back over ROTATING input sets whose total size exceeds the 126 MB L2 several times over, so every
call streams cold data, inside ONE event pair; the figure is the mean device time per call (sustained, no launch latency).

    [PAA_LIBPAA=variant.so] [PAA_STFT_HW=1] python tools/ab_time.py [case ...]      cases: bench.SWEEP_CASES names, compose
"""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import paa_b200  # noqa: E402
from paa_b200.core import iso  # noqa: E402
from paa_b200.core.compose import compose_clamp  # noqa: E402
from paa_b200.training_utils import build as pbuild  # noqa: E402

dev = torch.device("cuda:0")
only = set(sys.argv[1:])
tag = os.path.basename(os.environ.get("PAA_LIBPAA", "libpaa.so")) + (" hw" if os.environ.get("PAA_STFT_HW") == "1" else "")
peak, _ = bench.measured_peak()
interp = iso.build_weight_interpolator()
FOOT = bench.SUSTAINED_FOOTPRINT          # bytes the rotating sets cover together


timed = bench.sustained_us      # back to back on rotating cold input sets inside one event pair


for name, B, sec, rows, sigma, opt in bench.SWEEP_CASES:
    if only and name not in only:
        continue
    norm = bench.case_norm(name)
    T = sec * bench.SR
    nbytes = bench.sweep_bytes(norm, opt, rows or B, B, T)
    nsets = max(2, min(24, math.ceil(FOOT / max(nbytes, 1))))
    args = bench.sweep_args(name, opt, dev)
    thr = pbuild.init_phon_threshold_tensor(args)
    sets = []
    for _ in range(nsets):
        clean, p, grad = bench.sweep_inputs(dev, B, sec, sigma, rows)
        optim = pbuild.create_optimizer(args, p)[0] if opt == "adam" else None
        sets.append((clean, p, grad, optim))
    calls = [(lambda s=s: paa_b200.step_and_project(s[1], s[2], s[0], args, interp, thr, optimizer=s[3])) for s in sets]
    us = timed(calls, max(2, 60 // nsets))
    rec = {"shape": f"{B}x{sec}s", "p_rows": sets[0][1].shape[0], "us": round(us, 2), "GB/s": round(nbytes / us / 1e3, 1),
           "frac": round(nbytes / us / 1e3 / peak, 4), "sets": nsets}
    print(tag, name, json.dumps(rec), flush=True)
    del sets, calls
    torch.cuda.empty_cache()

if not only or "compose" in only:
    for B in (32, 128):
        T = bench.SECONDS * bench.SR
        for rows in (1, B):
            fb, bb = 8 * B * T + 4 * rows * T, 8 * B * T + 8 * rows * T
            nsets = max(2, min(24, math.ceil(FOOT / fb)))
            g = torch.Generator(device=dev).manual_seed(7)
            sets = []
            for _ in range(nsets):
                clean = (torch.rand(B, T, generator=g, device=dev) * 2 - 1) * 0.9
                w = torch.randn(B, T, generator=g, device=dev)
                p = (torch.randn(rows, T, generator=g, device=dev) * 0.3).requires_grad_(True)
                sets.append((clean, w, p))
            f_us = timed([(lambda s=s: compose_clamp(s[0], s[2])) for s in sets], max(2, 60 // nsets))
            xs = [compose_clamp(s[0], s[2]) for s in sets]
            b_us = timed([(lambda s=s, x=x: torch.autograd.grad(x, s[2], s[1], retain_graph=True)) for s, x in zip(sets, xs)],
                         max(2, 60 // nsets))
            kind = "universal" if rows == 1 else "per_utterance"
            print(tag, f"compose_fwd_{kind}_{B}x10s", json.dumps({"us": round(f_us, 2), "GB/s": round(fb / f_us / 1e3, 1), "frac": round(fb / f_us / 1e3 / peak, 4)}))
            print(tag, f"compose_bwd_{kind}_{B}x10s", json.dumps({"us": round(b_us, 2), "GB/s": round(bb / b_us / 1e3, 1), "frac": round(bb / b_us / 1e3 / peak, 4)}), flush=True)
            del sets, xs
            torch.cuda.empty_cache()
