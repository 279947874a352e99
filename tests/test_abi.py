"""CPU-side checks of the C ABI: libpaa.so loads, exports every symbol include/paa.h declares, and its
host-only entry points (ISO-226 tables, WER counters, argument validation) agree with the reference's
golden vectors.  No kernel is launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden

import paa_b200
from paa_b200 import paa_lib as L
from paa_b200.core import iso


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "paa.h")).read()
    declared = set(re.findall(r"\b(paa_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(L.EXPORTS)
    lib = C.CDLL(L.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert L.lib.paa_version() >= 100
    assert L.lib.paa_status_string(0) == b"ok"


def test_iso_tables_against_reference():
    g = load_golden("iso_tables")
    for i, ph in enumerate(g["phons"]):
        got = L.iso226_spl(float(ph), g["freqs"])
        assert np.abs(got - g["spl"][i]).max() < 1e-9
    ph, fk, w = L.weight_grid()
    assert np.array_equal(ph, g["grid_phon"]) and np.array_equal(fk, g["grid_freq"])
    assert np.abs(w - g["grid_w"]).max() < 1e-12
    it = iso.build_weight_interpolator()
    assert np.abs(it(g["query"]) - g["query_w"]).max() < 1e-12
    for n_fft in (512, 1024):
        for phon in (20.0, 35.5):
            got = L.spl_thresh(n_fft, 16000, phon)
            assert np.abs(got - g[f"thr_{n_fft}_{phon}"]).max() < 1e-3      # phon tolerance of north_star
            assert np.array_equal(got, g[f"thr_{n_fft}_{phon}"])


def test_iso_class_mirrors_reference_behaviour():
    assert abs(iso.ISO226(20)(np.array([1000.0]))[0] - 20.00517) < 1e-5
    assert iso.ISO226(40)(np.array([100]))[0] == 64                      # integer in, truncated out
    for bad in (-1, 90.5):
        with pytest.raises(ValueError):
            iso.ISO226(bad)
    for bad in (19.0, 20001.0):
        with pytest.raises(ValueError):
            iso.ISO226(40)(np.array([bad]))
    f, p, spl = iso.compute_iso226_weight_matrix()
    assert spl.shape == (10, 30) and abs(spl.max() - 123.70539502730307) < 1e-9
    w = iso.perceptual_weight(spl)
    np.testing.assert_allclose(w[:, 17], [1, 0.844866, 0.70272, 0.573687, 0.457747, 0.354888, 0.265105, 0.188394,
                                          0.124754, 0.074184], atol=1e-6)


def test_status_codes_without_gpu():
    out = np.empty(1)
    f = np.array([10.0])
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))           # noqa: E731
    assert L.lib.paa_iso226_spl(40.0, dp(f), 1, dp(out)) == L.ERR_RANGE
    assert L.lib.paa_iso226_spl(95.0, dp(np.array([100.0])), 1, dp(out)) == L.ERR_RANGE
    assert L.lib.paa_iso226_spl(40.0, None, 1, dp(out)) == L.ERR_NULL
    h = C.c_void_p()
    assert L.lib.paa_create(0, 768, 256, 16000, C.byref(h)) == L.ERR_UNSUPPORTED
    assert L.lib.paa_create(0, 1024, 300, 16000, C.byref(h)) == L.ERR_UNSUPPORTED
    assert L.lib.paa_create(0, 1024, 256, 16000, None) == L.ERR_NULL
    assert L.lib.paa_destroy(None) == L.OK
    assert L.lib.paa_scratch_bytes(None, 1, 1) == 0
    with pytest.raises(ValueError):
        L.check(L.ERR_NEED_CLEAN)
    with pytest.raises(ValueError):
        L.check(L.ERR_RANGE)
    with pytest.raises(RuntimeError):
        L.check(L.ERR_NOLA)
    with pytest.raises(L.PaaError):
        L.check(L.ERR_ALIAS)


def test_wer_counts():
    assert L.wer_counts(["a b c"], ["a b c"]) == (0, 3)
    assert L.wer_counts(["a b c", "hello world"], ["a x c d", ""]) == (4, 5)
    assert L.wer_counts(["  spaced   out  "], ["spaced out"]) == (0, 2)
    assert L.wer_counts([], []) == (0, 0)
    # known answers worked by hand (jiwer / HF evaluate are not in the image): word-level Levenshtein, summed over pairs
    assert L.wer_counts(["the cat sat on the mat"], ["the cat sit on mat"]) == (2, 6)         # 1 substitution + 1 deletion
    assert L.wer_counts(["a b c d"], ["b c d e"]) == (2, 4)                                  # 1 deletion + 1 insertion
    assert L.wer_counts(["hello"], [""]) == (1, 1)                                           # everything deleted
    assert L.wer_counts(["one two three"], ["one two three four five"]) == (2, 3)            # insertions count against the reference length
    assert L.wer_counts(["delete delete delete delete delete"] * 2, ["delete", "x y z"]) == (4 + 5, 10)
    from paa_b200.core.loss_helpers import WerMetric
    m = WerMetric()
    assert m.compute(predictions=["a"], references=["a b"]) == 0.5
    assert (m.errors, m.words) == (1, 2)


def test_cpu_tensors_are_refused():
    import torch
    args = paa_b200.training_utils.parser.create_arg_parser().parse_args(["--norm_type", "l2"])
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        paa_b200.perturbation_constraint(torch.zeros(1, 4096), None, args, None, None)


def test_checkpoint_formats_roundtrip(tmp_path):
    """perturbation.pt / results.json as the reference writes them (save.py:155-156, :226-256)."""
    import json
    import torch
    from paa_b200.training_utils import save
    p = torch.randn(1, 1000)
    path = tmp_path / "perturbation.pt"
    save.save_pert(p.requires_grad_(True), str(path))
    back = torch.load(str(path))                       # plain torch.load, as build.py:294-299 does
    assert back.dtype == torch.float32 and back.shape == (1, 1000) and not back.requires_grad
    assert torch.equal(back, p.detach()) and torch.equal(save.load_pert(str(path), "cpu"), p.detach())
    r = save.save_json_results(str(tmp_path), "snr", 40, epoch=3, test_loss_clean=2.0, test_loss_perturbed=5.0,
                               skipped=None, per_split={"a": 1.23456})
    on_disk = json.load(open(tmp_path / "results.json"))
    assert on_disk == r and r["perturbation_efficiency"] == 2.5 and "skipped" not in r and r["per_split"] == {"a": 1.2346}


def test_host_tables_against_scipy_directly():
    """Independent of the golden files: libpaa's PCHIP contour and bilinear lookup against scipy's own classes on
    random queries (the reference builds its tables with exactly these two, iso.py:113-124 and :261-266)."""
    interpolate = pytest.importorskip("scipy.interpolate")
    rng = np.random.default_rng(3)
    ph, fk, w = L.weight_grid()
    rgi = interpolate.RegularGridInterpolator((ph, fk), w, bounds_error=False, fill_value=1.0)
    q = np.stack([rng.uniform(-20, 110, 5000), np.exp(rng.uniform(np.log(2), np.log(40000), 5000))], 1)
    q[:50, 0] = rng.choice(ph, 50)                       # exactly on phon knots
    q[50:100, 1] = rng.choice(fk, 50)                    # exactly on frequency knots
    assert np.abs(L.interp2(ph, fk, w, 1.0, q) - rgi(q)).max() < 1e-12
    # the contour itself: PCHIP of the three ISO tables with the wrapped 20 kHz knot, then the closed form
    from oracle import paa_oracle as orc
    f = np.exp(rng.uniform(np.log(20), np.log(20000), 2000))
    xk = np.concatenate([orc.ISO_BAND_HZ, [20000.0]])
    al, lu, tf = (interpolate.PchipInterpolator(xk, np.concatenate([t, t[:1]]))(f) for t in (orc.ISO_ALPHA, orc.ISO_LU, orc.ISO_TF))
    for phon in (0.0, 13.7, 55.0, 90.0):
        a = 0.00447 * (10.0 ** (0.025 * phon) - 1.15)
        want = (10.0 / al) * np.log10(a + (0.4 * 10.0 ** ((tf + lu) / 10.0 - 9.0)) ** al) - lu + 94.0
        assert np.abs(L.iso226_spl(phon, f) - want).max() < 1e-9
        assert np.abs(orc.iso226_spl(phon, f) - want).max() < 1e-9


def test_ctypes_structs_match_the_header(tmp_path):
    """struct paa_step / paa_parts as the C compiler lays them out against the ctypes mirrors in paa_lib.py."""
    import ctypes
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    from paa_b200 import paa_lib as L
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "paa.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu %d\\n", sizeof(paa_step), offsetof(paa_step, grad), '
                   'offsetof(paa_step, adam_t), offsetof(paa_step, eps), offsetof(paa_step, parts), sizeof(paa_parts), '
                   'offsetof(paa_parts, clean_stats), offsetof(paa_parts, clean_numel), PAA_MAX_PARTS); return 0; }\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    want = [ctypes.sizeof(L.Step), L.Step.grad.offset, L.Step.adam_t.offset, L.Step.eps.offset, L.Step.parts.offset,
            ctypes.sizeof(L.Parts), L.Parts.clean_stats.offset, L.Parts.clean_numel.offset, L.MAX_PARTS]
    assert got == want
