"""ISO-226 equal-loudness tables with the reference's names (src/core/iso.py:34-266).  The numbers
come from libpaa.so's host code (PCHIP over the 29 third-octave bands plus the wrapped 20 kHz knot,
closed-form contour), not from scipy."""
from types import MappingProxyType
from typing import Union

import numpy as np

try:
    from .. import paa_lib as L
except ImportError:
    import paa_lib as L


class ISO226:
    """``ISO226(phon)(frequencies_hz)`` -> SPL in dB needed at each frequency for that loudness level."""

    reference = MappingProxyType({
        "frequencies": (20.0, 25.0, 31.5, 40.0, 50.0, 63.0, 80.0, 100.0, 125.0, 160.0, 200.0, 250.0, 315.0, 400.0,
                        500.0, 630.0, 800.0, 1000.0, 1250.0, 1600.0, 2000.0, 2500.0, 3150.0, 4000.0, 5000.0, 6300.0,
                        8000.0, 10000.0, 12500.0),
    })

    def __init__(self, phon: Union[int, float]) -> None:
        if phon < 0 or phon > 90:
            raise ValueError("Phon must be in range [0, 90]")
        self._phon = phon

    def __call__(self, frequencies) -> np.ndarray:
        frequencies = np.asarray(frequencies)
        if np.any(frequencies < 20.0) or np.any(frequencies > 20000.0):
            raise ValueError("Frequency must be in [20, 20000] Hz")
        out = np.zeros_like(frequencies)          # keeps the caller's dtype, as the reference does (iso.py:156)
        out[...] = L.iso226_spl(float(self._phon), frequencies.astype(np.float64))
        return out


def compute_iso226_weight_matrix():
    """(freqs[30], phons[10], spl[10,30]) -- iso.py:176-199."""
    freqs = np.array(ISO226.reference["frequencies"] + (20000.0,))
    phons = np.arange(0, 100, 10)
    spl = np.array([ISO226(ph)(freqs) for ph in phons])
    return freqs, phons, spl


def perceptual_weight(spl_matrix: np.ndarray) -> np.ndarray:
    """clip((1 - spl/max)^2, 0, 1) -- iso.py:202-235."""
    return np.clip((1 - (spl_matrix / spl_matrix.max())) ** 2, 0, 1)


class WeightInterpolator:
    """What the reference gets from scipy's RegularGridInterpolator(bounds_error=False, fill_value=1.0):
    a callable over (phon, freq) points that also exposes ``grid`` and ``values``.  The device kernels
    read ``grid``/``values``; calling it evaluates on the host (known-answer tests, plots)."""
    method = "linear"
    bounds_error = False

    def __init__(self, points, values, fill_value=1.0):
        self.grid = tuple(np.asarray(g, dtype=np.float64) for g in points)
        self.values = np.asarray(values, dtype=np.float64)
        self.fill_value = fill_value

    def __call__(self, xi):
        xi = np.asarray(xi, dtype=np.float64)
        return L.interp2(self.grid[0], self.grid[1], self.values, self.fill_value, xi).reshape(xi.shape[:-1])


def build_weight_interpolator():
    """iso.py:238-266."""
    phons, freqs, w = L.weight_grid()
    return WeightInterpolator((phons, freqs), w, fill_value=1.0)
