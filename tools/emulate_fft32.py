"""Index-math check of the half-warp real FFT of paa_fft32.cuh (n_fft 1024: N = 512 complex points, 16 lanes x 32 points).

Everything the CUDA code does with compile-time register names is replayed here with numpy arrays indexed the same way:
  forward   DFT-32 over q (thread l holds z[l + 16 q]) -> twiddle W_512^{l k1} -> exchange -> DFT-16 over l
            (thread lam holds k1 in {j0, j1}, outputs k = k1 + 32 k2)
  middle    conjugate pairs (k, N-k) inside one thread (slot s: za[s], zb[15-s]); lane 0 permuted onto the same slots
  inverse   IDFT-16 over k2 -> exchange -> twiddle -> IDFT-32 over k1, overlap-add weights w and 1-w
and compared with numpy.fft.rfft / irfft.   python tools/emulate_fft32.py
"""
import numpy as np

N, NFFT = 512, 1024


def W(n, d, sign=-1):
    return np.exp(sign * 2j * np.pi * n / d)


def dft4(v, sign):          # natural order in / out
    a0, a1, a2, a3 = v
    t0, t1, t2, t3 = a0 + a2, a0 - a2, a1 + a3, a1 - a3
    r = 1j * sign           # forward: -i
    return [t0 + t2, t1 + r * t3, t0 - t2, t1 - r * t3]


def dft8(v, sign):
    return list(np.array([sum(v[n] * W(n * k, 8, sign) for n in range(8)) for k in range(8)]))


def dft32(v, sign):
    """q = 8a + b, k = c + 4d: dft4 over a -> twiddle W_32^{bc} -> dft8 over b; natural order out."""
    v = list(v)
    for b in range(8):
        y = dft4([v[b], v[8 + b], v[16 + b], v[24 + b]], sign)
        for c in range(4):
            v[8 * c + b] = y[c] * W(b * c, 32, sign)
    out = [0] * 32
    for c in range(4):
        x = dft8(v[8 * c:8 * c + 8], sign)
        for d in range(8):
            out[c + 4 * d] = x[d]
    return out


def dft16(v, sign):
    """n = 4a + b, k = c + 4d: dft4 over a -> twiddle W_16^{bc} -> dft4 over b."""
    v = list(v)
    for b in range(4):
        y = dft4([v[b], v[4 + b], v[8 + b], v[12 + b]], sign)
        for c in range(4):
            v[4 * c + b] = y[c] * W(b * c, 16, sign)
    out = [0] * 16
    for c in range(4):
        x = dft4(v[4 * c:4 * c + 4], sign)
        for d in range(4):
            out[c + 4 * d] = x[d]
    return out


def j0(lam):
    return lam


def j1(lam):
    return 32 - lam if lam else 16


def forward(z):
    """z[512] -> per-thread registers za[lam][k2] = Z[j0 + 32 k2], zb[lam][k2] = Z[j1 + 32 k2]."""
    buf = np.zeros((32, 16), complex)                     # buf[k1][l] (the CUDA code pads the row to 17)
    for l in range(16):
        a = dft32([z[l + 16 * q] for q in range(32)], -1)
        for k1 in range(32):
            buf[k1][l] = a[k1] * W(l * k1, 512, -1)
    za = np.zeros((16, 16), complex)
    zb = np.zeros((16, 16), complex)
    for lam in range(16):
        za[lam] = dft16(list(buf[j0(lam)]), -1)
        zb[lam] = dft16(list(buf[j1(lam)]), -1)
    return za, zb


def inverse(za, zb):
    buf = np.zeros((32, 16), complex)
    for lam in range(16):
        buf[j0(lam)] = dft16(list(za[lam]), +1)
        buf[j1(lam)] = dft16(list(zb[lam]), +1)
    z = np.zeros(512, complex)
    for l in range(16):
        v = dft32([buf[k1][l] * W(l * k1, 512, +1) for k1 in range(32)], +1)
        for q in range(32):
            z[l + 16 * q] = v[q]
    return z


def lane0_in(za, zb):
    a, b = za.copy(), zb.copy()
    zb[15] = a[8]
    for s in range(1, 8):
        zb[15 - s] = a[16 - s]
    for i in range(8):
        za[8 + i] = b[i]
        zb[7 - i] = b[15 - i]


def lane0_out(za, zb):
    a, b = za.copy(), zb.copy()
    za[8] = b[15]
    for s in range(1, 8):
        za[16 - s] = b[15 - s]
    for i in range(8):
        zb[i] = a[8 + i]
        zb[15 - i] = b[7 - i]


def middle(za, zb, gain):
    """split -> X'[k] = gain[k] X[k] / 1024 -> merge, on un-halved Z (X2 = 2 X).  Returns the spectrum seen (for checking)."""
    X = np.zeros(N + 1, complex)
    kappa = 1.0 / 2048.0
    for lam in range(16):
        a, b = za[lam], zb[lam]
        l0 = lam == 0
        if l0:
            lane0_in(a, b)
        half_in = b[15]
        for s in range(16):
            if l0 and s >= 8:
                k = 32 * s - 240
            else:
                k = lam + 32 * s
            kn = N - k
            w = np.exp(2j * np.pi * k / NFFT)              # (cos, sin) of 2 pi k / n_fft
            dc = l0 and s == 0
            zA, zB = a[s], (a[s] if dc else b[15 - s])
            E, D = zA + np.conj(zB), zA - np.conj(zB)
            T = -1j * D * np.conj(w)
            X2, Yc2 = E + T, E - T                          # 2 X[k], 2 conj(X[N-k])
            X[k] = X2 / 2
            X[kn] = np.conj(Yc2) / 2
            Xo, Yo = X2 * gain[k] * kappa, Yc2 * gain[kn] * kappa
            if dc:
                Xo, Yo = Xo.real + 0j, Yo.real + 0j
            A, B = Xo + Yo, Xo - Yo
            p = w * B
            a[s] = A + 1j * p
            if not dc:
                b[15 - s] = np.conj(A) + (p.imag + 1j * p.real)      # conj(A) + swap(p)
        if l0:                                              # the self-paired bin N/2
            Xh2 = 2 * np.conj(half_in)
            X[N // 2] = Xh2 / 2
            b[15] = 2 * np.conj(Xh2 * gain[N // 2] * kappa)
            lane0_out(a, b)
    return X


def main():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(NFFT)
    z = x[0::2] + 1j * x[1::2]
    za, zb = forward(z)
    Z = np.fft.fft(z)
    for lam in range(16):
        for k2 in range(16):
            assert abs(za[lam][k2] - Z[j0(lam) + 32 * k2]) < 1e-9
            assert abs(zb[lam][k2] - Z[j1(lam) + 32 * k2]) < 1e-9
    gain = rng.uniform(0.2, 1.5, N + 1)
    X = middle(za, zb, gain)
    Xref = np.fft.rfft(x)
    assert np.abs(X - Xref).max() < 1e-9, np.abs(X - Xref).max()
    zi = inverse(za, zb)
    y = np.empty(NFFT)
    y[0::2], y[1::2] = zi.real, zi.imag
    Xg = Xref * gain
    Xg[0], Xg[N] = Xg[0].real, Xg[N].real
    yref = np.fft.irfft(Xg, NFFT)
    assert np.abs(y - yref).max() < 1e-9, np.abs(y - yref).max()
    # exchange layout: row stride 17 makes the column accesses of the 16 lanes hit 16 distinct 8-byte slots mod 16
    for l in range(16):
        assert len({(17 * j0(lam) + l) % 16 for lam in range(16)}) == 16
        assert len({(17 * j1(lam) + l) % 16 for lam in range(16)}) == 16
    # window: w[n + 512] = 1 - w[n]
    w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(NFFT) / NFFT)
    assert np.abs(w[512:] - (1 - w[:512])).max() < 1e-15
    print("emulate_fft32: forward, paired middle (incl. lane 0), inverse, layout: ok")


if __name__ == "__main__":
    main()
