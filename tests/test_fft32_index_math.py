"""CPU check of the half-warp FFT's index math (paa_fft32.cuh): tools/emulate_fft32.py replays the decomposition, the
paired spectral middle incl. lane 0's permutation, the inverse and the exchange-buffer layout with numpy against
numpy.fft; the CUDA kernel itself is covered on the GPU by test_half_warp_kernel_variant_passes_the_same_parity_tests."""
import importlib.util
import os


def test_emulated_half_warp_fft_matches_numpy(capsys):
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "emulate_fft32.py")
    spec = importlib.util.spec_from_file_location("emulate_fft32", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.main()
    assert "ok" in capsys.readouterr().out
