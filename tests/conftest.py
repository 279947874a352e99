import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def golden_manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as fh:
        return json.load(fh)


def rel_max(a, b):
    """max|a-b| / max|b|  -- the parity metric of SURVEY.md section 8(d)."""
    a = torch.as_tensor(a).double().reshape(-1)
    b = torch.as_tensor(b).double().reshape(-1)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def rel_l2(a, b):
    a = torch.as_tensor(a).double().reshape(-1)
    b = torch.as_tensor(b).double().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def hyper_from_manifest(entry, **extra):
    from oracle import paa_oracle as orc
    kw = dict(entry["defaults"])
    for k in ("norm_type", "lr", "n_fft", "hop_length", "sr"):
        kw[k] = entry[k]
    for k in ("l2_size", "linf_size", "snr_db", "tv_epsilon", "fm_epsilon", "min_freq_attack", "max_freq_attack"):
        if k in entry:
            kw[k] = entry[k]
    kw["win_length"] = kw["n_fft"]
    kw.update(extra)
    return orc.Hyper(**kw)
