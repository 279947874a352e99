"""Profiling driver (run under ncu on the GPU box): one attack step of bench.py's workload followed by one
step+projection call per norm_type at BASELINE.json's shapes.  Prints per-call CUDA-event times."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import paa_b200  # noqa: E402
from paa_b200.core import iso, loss_helpers  # noqa: E402
from paa_b200.training_utils import build as pbuild, parser as pparser  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--no-model", action="store_true")
ap.add_argument("--norms", default="linf,l2,snr,tv,min_max_freqs,max_phon,fletcher_munson")
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--compose", action="store_true", help="also one compose + clamp forward / backward per p shape at 128 x 10 s")
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
interp = iso.build_weight_interpolator()

if not a.no_model:
    T = bench.SECONDS * bench.SR
    args = pparser.create_arg_parser().parse_args(["--norm_type", "snr", "--snr_db", "40", "--optimizer_type", "pgd",
                                                   "--attack_mode", "targeted"])
    args.device = str(dev)
    model = bench.build_model(dev)
    for q in model.parameters():
        q.requires_grad_(False)
    clean, p0 = bench.synth(0, bench.C1_BATCH, T, dev, bench.C1_BATCH)
    clean = clean.to(dev)
    labels = loss_helpers.encode_labels(["delete delete delete delete delete"] * bench.C1_BATCH, dev)
    p = paa_b200.perturbation_constraint(p0.to(dev), clean, args, None, None)
    for it in range(2):
        p = p.detach().requires_grad_(True)
        out = model(input_values=(clean + p).clamp_(-1, 1), labels=labels)
        (-out.loss).backward()
        p = paa_b200.step_and_project(p.detach(), p.grad, clean, args, None, None)
    torch.cuda.synchronize()
    del model, clean, p, out

shapes = {"linf": (4, 5, 1e-3), "snr": (32, 10, 0.01), "fletcher_munson": (64, 15, 0.1), "max_phon": (64, 15, 0.03),
          "tv": (128, 10, 0.01), "min_max_freqs": (128, 10, 0.01), "l2": (512, 10, 0.01)}
for norm in a.norms.split(","):
    B, sec, sigma = shapes[norm]
    T = sec * bench.SR
    g = torch.Generator(device=dev).manual_seed(1234)
    clean = (torch.rand(B, T, generator=g, device=dev) * 2 - 1) * 0.1
    p = torch.randn(B, T, generator=g, device=dev) * sigma
    grad = torch.randn(B, T, generator=g, device=dev)
    args = pparser.create_arg_parser().parse_args(["--norm_type", norm, "--optimizer_type", "pgd", "--snr_db", "40"])
    args.device = str(dev)
    thr = pbuild.init_phon_threshold_tensor(args)
    for r in range(a.reps + 1):
        torch.cuda._sleep(400_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        paa_b200.step_and_project(p, grad, clean, args, interp, thr)
        e1.record()
        torch.cuda.synchronize()
    print(f"{norm:16s} {B}x{sec}s  {e0.elapsed_time(e1) * 1e3:9.1f} us", flush=True)
    del clean, p, grad

if a.compose:
    from paa_b200.core.compose import compose_clamp  # noqa: E402
    B, T = 128, 10 * bench.SR
    g = torch.Generator(device=dev).manual_seed(7)
    clean = (torch.rand(B, T, generator=g, device=dev) * 2 - 1) * 0.9
    w = torch.randn(B, T, generator=g, device=dev)
    for rows in (1, B):
        p = (torch.randn(rows, T, generator=g, device=dev) * 0.3).requires_grad_(True)
        for r in range(2):
            x = compose_clamp(clean, p)
            torch.autograd.grad(x, p, w)
        torch.cuda.synchronize()
        print(f"compose fwd+bwd  p_rows {rows}", flush=True)
