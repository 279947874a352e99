"""Small driver for compute-sanitizer: every norm_type once (PGD-fused and Adam), tiny shapes, plus the un-fused
STFT/ISTFT and spectrum functions."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import paa_b200  # noqa: E402
from paa_b200.core import fourier_transforms as ft, iso, projections as pj  # noqa: E402
from paa_b200.training_utils import build as pbuild, parser as pparser  # noqa: E402

dev = torch.device("cuda:0")
interp = iso.build_weight_interpolator()
g = torch.Generator(device=dev).manual_seed(0)
for n_fft, hop in ((1024, 256), (512, 256), (512, 128)):
    for norm in ("linf", "l2", "snr", "tv", "min_max_freqs", "max_phon", "fletcher_munson"):
        for B, T in ((3, 9000), (1, 4999)):
            clean = (torch.rand(B, T, generator=g, device=dev) * 2 - 1) * 0.1
            p = torch.randn(B, T, generator=g, device=dev) * 0.05
            grad = torch.randn(B, T, generator=g, device=dev)
            args = pparser.create_arg_parser().parse_args(
                ["--norm_type", norm, "--optimizer_type", "pgd", "--snr_db", "40", "--n_fft", str(n_fft),
                 "--hop_length", str(hop), "--win_length", str(n_fft)])
            args.device = str(dev)
            thr = pbuild.init_phon_threshold_tensor(args)
            out = paa_b200.step_and_project(p, grad, clean, args, interp, thr)
            out2 = paa_b200.perturbation_constraint(p, None if norm in ("snr", "tv") else None, args, interp, thr) \
                if norm not in ("snr", "tv") else paa_b200.perturbation_constraint(p, clean, args, interp, thr)
            args.optimizer_type = "adam"
            pa = torch.nn.Parameter(p.clone())
            opt, _ = pbuild.create_optimizer(args, pa)
            pa.data = paa_b200.step_and_project(pa.data, grad, clean, args, interp, thr, optimizer=opt)
            assert torch.isfinite(out).all() and torch.isfinite(out2).all() and torch.isfinite(pa.data).all()
    x = torch.randn(2, 6000, generator=g, device=dev) * 0.05
    S = ft.compute_stft(x, args)
    y = ft.compute_istft(S, args)
    pj.project_min_max_freqs(args, S, 300.0, 3400.0)
    pj.project_phon_level(S, args, thr)
    pj.project_fm_norm(S, args, interp)
torch.cuda.synchronize()
print("sanitize_case ok")
