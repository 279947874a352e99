"""Numpy emulation of csrc/paa_fft.cuh + spectral_middle (index logic only), checked against numpy's rfft/irfft."""
import numpy as np

PLANS = {1024: (512, (8, 8, 8)), 512: (256, (4, 8, 8))}

def dft(v, sign):
    R = len(v); k = np.arange(R)
    W = np.exp(sign * 2j * np.pi * np.outer(k, k) / R)
    return W @ v

def fft_warp(N, radices, first, sign):
    buf = np.zeros(N, complex)
    Ns = 1
    out = np.zeros(N, complex)
    for si, R in enumerate(radices):
        NB = N // R // 32
        vals = {}
        for lane in range(32):
            for b in range(NB):
                j = lane + 32 * b
                v = np.array([(first(j + r * (N // R)) if si == 0 else buf[j + r * (N // R)]) for r in range(R)])
                k = j % Ns
                tw = np.exp(sign * 2j * np.pi * k * np.arange(R) / (Ns * R))
                vals[(lane, b)] = dft(v * tw, sign)
        for (lane, b), v in vals.items():
            j = lane + 32 * b
            base = (j // Ns) * Ns * R + (j % Ns)
            for r in range(R):
                if si == len(radices) - 1:
                    assert base == j
                    out[j + r * Ns] = v[r]
                else:
                    buf[base + r * Ns] = v[r]
        Ns *= R
    return out

def middle(n_fft, Z, op):
    N = n_fft // 2
    post = np.exp(2j * np.pi * np.arange(N // 2 + 1) / n_fft)   # (cos, sin)
    Zp = np.zeros(N, complex); X_all = np.zeros(N + 1, complex)
    def pair(k):
        kn = N - k
        za, zb = Z[k], Z[kn & (N - 1)]
        w = post[k]
        er, ei = 0.5 * (za.real + zb.real), 0.5 * (za.imag - zb.imag)
        o_r, o_i = 0.5 * (za.imag + zb.imag), -0.5 * (za.real - zb.real)
        tr, ti = w.real * o_r + w.imag * o_i, w.real * o_i - w.imag * o_r
        X = complex(er + tr, ei + ti); Y = complex(er - tr, -(ei - ti))
        X_all[k] = X
        if kn != k: X_all[kn] = Y
        X = op(k, X); Y = op(kn, Y) if kn != k else X
        if k == 0: X = complex(X.real, 0); Y = complex(Y.real, 0)
        ar, ai, br, bi = X.real + Y.real, X.imag - Y.imag, X.real - Y.real, X.imag + Y.imag
        pr, pi = w.real * br - w.imag * bi, w.real * bi + w.imag * br
        Zp[k] = complex(ar - pi, ai + pr) / n_fft
        if k != 0: Zp[kn] = complex(ar + pi, pr - ai) / n_fft
    for i in range(N // 64):
        for lane in range(32): pair(lane + 32 * i)
    pair(N // 2)
    return X_all, Zp

for n_fft, (N, rad) in PLANS.items():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(n_fft)
    z = x[0::2] + 1j * x[1::2]
    Z = fft_warp(N, rad, lambda m: z[m], -1)
    assert np.allclose(Z, np.fft.fft(z)), "complex fft"
    gain = rng.uniform(0.2, 2, N + 1) * np.exp(1j * rng.uniform(-1, 1, N + 1))
    X, Zp = middle(n_fft, Z, lambda k, v: v * gain[k])
    assert np.allclose(X, np.fft.rfft(x)), "split"
    zz = fft_warp(N, rad, lambda m: Zp[m], +1)
    y = np.empty(n_fft); y[0::2] = zz.real; y[1::2] = zz.imag
    Xg = np.fft.rfft(x) * gain
    assert np.allclose(y, np.fft.irfft(Xg, n_fft)), "merge+inverse"
    print(n_fft, "ok")
