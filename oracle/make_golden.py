"""Generate tests/golden/*.npz by running the REFERENCE ITSELF (imported from
/root/reference/src, which exists only in the build container).  Test infrastructure.

    python oracle/make_golden.py            # rewrites tests/golden/

The reference imports matplotlib only for debug plots; it is stubbed.
The fixtures are what pins oracle/paa_oracle.py (tests/test_oracle_golden.py) and, on the
GPU box where /root/reference does not exist, the CUDA path (tests/test_gpu_golden.py).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import types

import numpy as np
import torch

REF_SRC = "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, REF_SRC)
    from core import fourier_transforms, iso, projections          # noqa
    from training_utils import build, parser, train               # noqa
    return types.SimpleNamespace(ft=fourier_transforms, iso=iso, proj=projections,
                                 build=build, parser=parser, train=train)


def ref_args(ref, **kw):
    a = ref.parser.create_arg_parser().parse_args([])
    a.device = "cpu"
    for k, v in kw.items():
        setattr(a, k, v)
    a.win_length = a.n_fft
    return a


def inputs(seed, rows, B, T, sigma, zero_frac=0.01):
    g = torch.Generator().manual_seed(seed)
    clean = (torch.rand(B, T, generator=g) * 2 - 1) * 0.1
    p = torch.randn(rows, T, generator=g) * sigma
    grad = torch.randn(rows, T, generator=g)
    grad[torch.rand(rows, T, generator=g) < zero_frac] = 0.0
    return clean, p, grad


def save(name, **arrs):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
                                 for k, v in arrs.items()})
    return path


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = import_reference()
    torch.set_num_threads(1)
    manifest = {}

    # ---- ISO-226 tables -------------------------------------------------------------
    fq = np.concatenate([np.array(ref.iso.ISO226.reference["frequencies"] + (20000.0,)),
                         np.geomspace(20.0, 20000.0, 61), [21.3, 999.9, 15999.0]])
    phons = [0.0, 7.5, 20.0, 35.5, 40.0, 60.0, 90.0]
    spl = np.stack([ref.iso.ISO226(ph)(fq) for ph in phons])
    interp = ref.iso.build_weight_interpolator()
    rng = np.random.default_rng(7)
    q = np.stack([rng.uniform(-15, 105, 400), np.exp(rng.uniform(np.log(5), np.log(30000), 400))], 1)
    q = np.concatenate([q, [[35, 1000], [63.2, 912], [35, 20], [90, 8000], [0, 20], [90, 20000],
                            [-5, 100], [95, 100], [50, 15.625], [10, 12500], [20, 12500.0001]]])
    thr = {}
    for n_fft in (512, 1024):
        for ph in (20.0, 35.5):
            a = ref_args(ref, n_fft=n_fft, max_phon_level=ph)
            thr[f"thr_{n_fft}_{ph}"] = ref.build.init_phon_threshold_tensor(a).reshape(-1)
    save("iso_tables", freqs=fq, phons=np.array(phons), spl=spl,
         grid_phon=interp.grid[0], grid_freq=interp.grid[1], grid_w=interp.values,
         query=q, query_w=interp(q), **thr)

    # ---- STFT / ISTFT ---------------------------------------------------------------
    for n_fft, hop, T in ((1024, 256, 3000), (1024, 256, 2560), (512, 256, 3000), (512, 128, 1500)):
        a = ref_args(ref, n_fft=n_fft, hop_length=hop)
        _, p, _ = inputs(100 + n_fft + T, 2, 2, T, 0.05)
        S = ref.ft.compute_stft(p, a)
        y = ref.ft.compute_istft(S, a)
        save(f"stft_{n_fft}_{hop}_{T}", x=p, spec=torch.view_as_real(S.contiguous()), y=y)

    # ---- projection only, PGD-fused and Adam-fused, every norm_type ------------------
    spl20 = {n: ref.build.init_phon_threshold_tensor(ref_args(ref, n_fft=n)) for n in (512, 1024)}
    cases = [
        # name, norm, hyper overrides, sigma, T
        ("linf", "linf", dict(linf_size=1e-4), 1e-3, 3000),
        ("l2_bind", "l2", dict(l2_size=0.05), 0.01, 3000),
        ("l2_free", "l2", dict(l2_size=50.0), 0.01, 3000),
        ("snr_bind", "snr", dict(snr_db=40.0), 0.01, 3000),
        ("snr_free", "snr", dict(snr_db=5.0), 0.01, 3000),
        ("tv_bind", "tv", dict(tv_epsilon=1e-3), 0.01, 3000),
        ("tv_free", "tv", dict(tv_epsilon=10.0), 0.01, 3000),
        ("mmf_default", "min_max_freqs", dict(), 0.01, 3000),
        ("mmf_tel", "min_max_freqs", dict(min_freq_attack=300.0, max_freq_attack=3400.0), 0.01, 2560),
        ("mmf_512", "min_max_freqs", dict(n_fft=512, min_freq_attack=300.0, max_freq_attack=3400.0), 0.01, 3000),
        ("phon_1024", "max_phon", dict(), 0.03, 3000),
        ("phon_1024_loud", "max_phon", dict(), 0.3, 2560),
        ("phon_512", "max_phon", dict(n_fft=512), 0.03, 3000),
        ("fm_bind", "fletcher_munson", dict(fm_epsilon=2.0), 0.1, 3000),
        ("fm_free", "fletcher_munson", dict(fm_epsilon=1e4), 0.1, 2560),
        ("fm_512", "fletcher_munson", dict(n_fft=512, fm_epsilon=1.0), 0.1, 3000),
    ]
    for i, (name, norm, over, sigma, T) in enumerate(cases):
        for rows in (1, 2):
            a = ref_args(ref, norm_type=norm, optimizer_type="pgd", lr=1e-4 if norm != "linf" else 5e-5, **over)
            clean, p, grad = inputs(1234 + i, rows, 2, T, sigma)
            thr_t = spl20[a.n_fft]
            with torch.no_grad():
                proj = ref.train.perturbation_constraint(p.clone(), clean, a, interp, thr_t)
                q_pgd = p.clone()
                q_pgd.add_(a.lr * grad.sign())                      # train.py:161
                pgd = ref.train.perturbation_constraint(q_pgd, clean, a, interp, thr_t)
                nocl = None
                if norm not in ("snr", "tv"):
                    nocl = ref.train.perturbation_constraint(p.clone(), None, a, interp, thr_t)
            # two Adam steps through torch.optim.Adam exactly as train.py:165-175 drives it
            pa = torch.nn.Parameter(p.clone())
            opt = torch.optim.Adam([pa], lr=a.lr)
            adam_out = []
            for s in range(2):
                opt.zero_grad(set_to_none=True)
                pa.grad = (grad * (1.0 if s == 0 else -0.5)).clone()
                opt.step()
                with torch.no_grad():
                    pa.data = ref.train.perturbation_constraint(pa.data, clean, a, interp, thr_t)
                adam_out.append(pa.data.clone())
            st = opt.state[pa]
            arrs = dict(clean=clean, p=p, grad=grad, proj=proj, pgd=pgd, adam1=adam_out[0], adam2=adam_out[1],
                        adam_m=st["exp_avg"], adam_v=st["exp_avg_sq"])
            if nocl is not None:
                arrs["proj_noclean"] = nocl
            save(f"norm_{name}_r{rows}", **arrs)
            manifest[f"norm_{name}_r{rows}"] = dict(
                norm_type=norm, rows=rows, T=T, lr=a.lr, n_fft=a.n_fft, hop_length=a.hop_length, sr=a.sr,
                **{k: v for k, v in over.items() if k != "n_fft"},
                defaults={k: getattr(a, k) for k in ("l2_size", "linf_size", "snr_db", "tv_epsilon", "fm_epsilon",
                                                     "min_freq_attack", "max_freq_attack", "max_phon_level",
                                                     "phon_reference_db")})

    # per-bin spectrum ops on their own (unfused public functions)
    a = ref_args(ref)
    _, p, _ = inputs(77, 2, 2, 3000, 0.05)
    S = ref.ft.compute_stft(p, a)
    save("spectrum_ops", x=p,
         mask=torch.view_as_real(ref.proj.project_min_max_freqs(a, S, 300.0, 3400.0).contiguous()),
         phon=torch.view_as_real(ref.proj.project_phon_level(S, a, spl20[1024]).contiguous()),
         fm_norm=ref.proj.compute_fm_weighted_norm_interp(S, interp, a),
         fm=torch.view_as_real(ref.proj.project_fm_norm(S, ref_args(ref, fm_epsilon=3.0), interp).contiguous()),
         level_db=20 * torch.log10(S.abs() + 1e-8))

    with open(os.path.join(OUT, "manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    print("wrote", len(os.listdir(OUT)), "files to", os.path.normpath(OUT))


if __name__ == "__main__":
    argparse.ArgumentParser(description=__doc__).parse_args()
    main()
