"""Brute-force search of shared-memory padding for the warp-per-frame Stockham FFT.
Complex points are stored as float2 (8 B); an LDS.64/STS.64 is served one half-warp at a
time, conflict-free when the 16 lanes hit 16 distinct 8-byte bank pairs (index mod 16).
pad(i) = i + sum_m c[m] * bit_m(i >> 4).  Reads are 16-aligned runs, so any such pad keeps
them conflict-free; the strided Stockham writes are what is scored here."""
import itertools, sys
import numpy as np

def stage_write_idx(N, R, Ns):
    nb = N // R // 32
    out = []
    for b in range(max(nb, 1)):
        lanes = np.arange(32)
        j = lanes + 32 * b
        valid = j < N // R
        for r in range(R):
            idx = (j // Ns) * Ns * R + (j % Ns) + r * Ns
            out.append((idx, valid))
    return out

def score(pad, accesses):
    worst, total = 1, 0
    for idx, valid in accesses:
        phys = pad(idx)
        for h in range(2):
            sel = valid[16*h:16*h+16]
            banks = phys[16*h:16*h+16][sel] % 16
            if banks.size == 0: continue
            m = np.bincount(banks, minlength=16).max()
            worst = max(worst, m); total += m
    return worst, total

def search(N, radices):
    acc = []
    Ns = 1
    for R in radices[:-1]:          # the last stage writes lane-consecutive runs
        acc += stage_write_idx(N, R, Ns); Ns *= R
    nbits = int(np.log2(N)) - 4
    best = None
    ideal = sum(2 for _ in acc)
    for c in itertools.product(range(16), repeat=nbits):
        c = np.array(c)
        def pad(i, c=c):
            hi = i >> 4
            g = 0
            for m in range(nbits):
                g = g + c[m] * ((hi >> m) & 1)
            return i + g
        w, t = score(pad, acc)
        key = (w, t, int(c.sum()))
        if best is None or key < best[0]:
            best = (key, tuple(int(x) for x in c))
            if w == 1 and False: break
    return best, ideal

if __name__ == "__main__":
    for N, rad in ((512, (8, 8, 8)), (256, (4, 8, 8)), (256, (8, 8, 4))):
        print(N, rad, search(N, rad))
