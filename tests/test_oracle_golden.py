"""Pins oracle/paa_oracle.py against outputs of the reference itself (tests/golden, made by
oracle/make_golden.py in the build container).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import golden_manifest, hyper_from_manifest, load_golden, rel_l2, rel_max
from oracle import paa_oracle as orc

MAN = golden_manifest()
TOL = 1e-5          # max|a-b|/max|b| and l2-relative, SURVEY.md section 8(d)


def test_iso226_contours_match_reference():
    g = load_golden("iso_tables")
    for i, ph in enumerate(g["phons"]):
        got = orc.iso226_spl(float(ph), g["freqs"])
        assert np.abs(got - g["spl"][i]).max() < 1e-9          # fp64 closed form + PCHIP


def test_iso226_known_answers():
    # SURVEY.md section 4 KATs (fp64)
    assert abs(orc.iso226_spl(20, np.array([1000.0]))[0] - 20.00517) < 1e-5
    assert abs(orc.iso226_spl(60, np.array([1000.0]))[0] - 60.011588) < 1e-5
    assert abs(orc.iso226_spl(40, np.array([100.0]))[0] - 64.371149) < 1e-5
    assert orc.iso226_spl(40, np.array([100]))[0] == 64          # integer in -> truncated out
    with pytest.raises(ValueError):
        orc.iso226_spl(91, np.array([100.0]))
    with pytest.raises(ValueError):
        orc.iso226_spl(40, np.array([19.9]))
    with pytest.raises(ValueError):
        orc.iso226_spl(40, np.array([20000.5]))


def test_weight_grid_and_interpolator():
    g = load_golden("iso_tables")
    ph, fk, w = orc.iso226_weight_grid()
    assert np.array_equal(ph, g["grid_phon"]) and np.array_equal(fk, g["grid_freq"])
    assert np.abs(w - g["grid_w"]).max() < 1e-12
    np.testing.assert_allclose(w[0, :5], [0.145296, 0.220482, 0.307362, 0.399318, 0.484127], atol=1e-6)
    it = orc.build_weight_interpolator()
    assert np.abs(it(g["query"]) - g["query_w"]).max() < 1e-12


@pytest.mark.parametrize("n_fft", [512, 1024])
@pytest.mark.parametrize("phon", [20.0, 35.5])
def test_phon_threshold(n_fft, phon):
    g = load_golden("iso_tables")
    got = orc.phon_threshold(n_fft, 16000, phon).reshape(-1).numpy()
    assert np.abs(got - g[f"thr_{n_fft}_{phon}"]).max() < 1e-3        # phon tolerance of north_star
    assert np.array_equal(got, g[f"thr_{n_fft}_{phon}"])               # and in fact bit-equal


@pytest.mark.parametrize("name", ["stft_1024_256_3000", "stft_1024_256_2560", "stft_512_256_3000",
                                  "stft_512_128_1500"])
def test_stft_istft(name):
    g = load_golden(name)
    _, n_fft, hop, _ = name.split("_")
    x = torch.from_numpy(g["x"])
    S = orc.stft(x, int(n_fft), int(hop))
    ref = torch.view_as_complex(torch.from_numpy(g["spec"]))
    assert S.shape == ref.shape
    assert rel_max(torch.view_as_real(S), torch.view_as_real(ref)) < 1e-6
    y = orc.istft(ref, int(n_fft), int(hop))
    assert y.shape == g["y"].shape
    assert rel_max(y, g["y"]) < 1e-6


@pytest.mark.parametrize("name", sorted(MAN))
def test_constraint_pgd_adam(name):
    e, g = MAN[name], load_golden(name)
    hp = hyper_from_manifest(e, optimizer_type="pgd")
    interp = orc.build_weight_interpolator()
    thr = orc.phon_threshold(hp.n_fft, hp.sr, hp.max_phon_level)
    clean, p, grad = (torch.from_numpy(g[k]) for k in ("clean", "p", "grad"))
    out = orc.constrain(p, clean, hp, interp, thr)
    assert rel_max(out, g["proj"]) < TOL and rel_l2(out, g["proj"]) < TOL
    out = orc.step_and_constrain(p, grad, clean, hp, interp, thr)
    assert rel_max(out, g["pgd"]) < TOL and rel_l2(out, g["pgd"]) < TOL
    if "proj_noclean" in g:
        out = orc.constrain(p, None, hp, interp, thr)
        assert out.shape == g["proj_noclean"].shape
        assert rel_max(out, g["proj_noclean"]) < TOL
    hp.optimizer_type = "adam"
    st = orc.AdamState(torch.zeros_like(p), torch.zeros_like(p))
    q = orc.step_and_constrain(p, grad, clean, hp, interp, thr, adam=st)
    assert rel_max(q, g["adam1"]) < TOL
    q = orc.step_and_constrain(q, grad * -0.5, clean, hp, interp, thr, adam=st)
    assert rel_max(q, g["adam2"]) < TOL and rel_l2(q, g["adam2"]) < TOL
    assert rel_max(st.m, g["adam_m"]) < 1e-6 and rel_max(st.v, g["adam_v"]) < 1e-6


def test_spectrum_ops():
    g = load_golden("spectrum_ops")
    hp = orc.Hyper()
    S = orc.stft(torch.from_numpy(g["x"]), 1024, 256)
    cplx = lambda k: torch.view_as_complex(torch.from_numpy(g[k]))     # noqa: E731
    m = S * orc.band_mask(1024, 16000, 300.0, 3400.0)
    assert rel_max(torch.view_as_real(m), g["mask"]) < 1e-6
    thr = orc.phon_threshold(1024, 16000, 20.0)
    ph = orc.project_phon(S, thr, 65.0)
    assert rel_max(torch.view_as_real(ph), g["phon"]) < 1e-6
    assert float((orc.phon_levels(S) - torch.from_numpy(g["level_db"])).abs().max()) < 1e-3
    it = orc.build_weight_interpolator()
    assert abs(float(orc.fm_weighted_norm(S, it, 1024, 16000)) / float(g["fm_norm"]) - 1) < 1e-6
    fm = orc.project_fm(S, it, 1024, 16000, 3.0)
    assert rel_max(torch.view_as_real(fm), g["fm"]) < 1e-6
    assert float(g["fm_norm"]) > 3.0      # the fixture exercises the scaling branch


def test_dispatch_errors():
    hp = orc.Hyper(norm_type="snr")
    with pytest.raises(ValueError):
        orc.constrain(torch.zeros(1, 4096), None, hp)
    hp.norm_type = "tv"
    with pytest.raises(ValueError):
        orc.constrain(torch.zeros(1, 4096), None, hp)
    hp.norm_type = "l7"
    with pytest.raises(ValueError):
        orc.constrain(torch.zeros(1, 4096), torch.zeros(1, 4096), hp)
    hp.norm_type, hp.optimizer_type = "l2", "sgd"
    with pytest.raises(NotImplementedError):
        orc.step_and_constrain(torch.zeros(1, 8), torch.zeros(1, 8), None, hp)


def test_edit_counts():
    assert orc.edit_counts(["a b c"], ["a b c"]) == (0, 3)
    assert orc.edit_counts(["a b c", "hello world"], ["a x c d", ""]) == (4, 5)
    assert orc.WerMetric().compute(predictions=["a"], references=["a b"]) == 0.5
