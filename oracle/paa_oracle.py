"""CPU oracle for the perturbation hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

A CPU restatement (torch-CPU fp32 for the tensor arithmetic, numpy fp64 for the
ISO-226 tables) of the reference's per-iteration path: the PGD / Adam step on the
waveform perturbation followed by the ``norm_type`` projection.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may
import this file; the product package must never do so (it fails loudly when the CUDA
library is missing instead).

Pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the *reference itself*, imported from
``/root/reference/src`` in the build container by ``oracle/make_golden.py`` and committed
as fixtures under ``tests/golden/``.  ``tests/test_oracle_golden.py`` replays them.

Every function cites the reference lines it restates (paths relative to the reference
repository root).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

# ----------------------------------------------------------------------------------------
# ISO-226 tables (src/core/iso.py:60-84).  29 third-octave bands, Hz.
# ----------------------------------------------------------------------------------------
ISO_BAND_HZ = np.array(
    [20, 25, 31.5, 40, 50, 63, 80, 100, 125, 160, 200, 250, 315, 400, 500, 630, 800, 1000,
     1250, 1600, 2000, 2500, 3150, 4000, 5000, 6300, 8000, 10000, 12500], dtype=np.float64)
ISO_ALPHA = np.array(
    [.532, .506, .480, .455, .432, .409, .387, .367, .349, .330, .315, .301, .288, .276, .267,
     .259, .253, .250, .246, .244, .243, .243, .243, .242, .242, .245, .254, .271, .301],
    dtype=np.float64)
ISO_LU = np.array(
    [-31.6, -27.2, -23.0, -19.1, -15.9, -13.0, -10.3, -8.1, -6.2, -4.5, -3.1, -2.0, -1.1, -0.4,
     0.0, 0.3, 0.5, 0.0, -2.7, -4.1, -1.0, 1.7, 2.5, 1.2, -2.1, -7.1, -11.2, -10.7, -3.1],
    dtype=np.float64)
ISO_TF = np.array(
    [78.5, 68.7, 59.5, 51.1, 44.0, 37.5, 31.5, 26.5, 22.1, 17.9, 14.4, 11.4, 8.6, 6.2, 4.4, 3.0,
     2.2, 2.4, 3.5, 1.7, -1.3, -4.2, -6.0, -5.4, -1.5, 6.0, 12.6, 13.9, 12.3], dtype=np.float64)

F_LO_HZ, F_HI_HZ = 20.0, 20000.0
PHON_KNOTS = np.arange(0.0, 100.0, 10.0)          # iso.py:189  (0,10,...,90)


def _knots30() -> np.ndarray:
    """Frequency axis with the extra 20 kHz knot (iso.py:104, :192)."""
    return np.concatenate([ISO_BAND_HZ, [F_HI_HZ]])


def _wrapped(table: np.ndarray) -> np.ndarray:
    """The 30th knot re-uses the *20 Hz* entry (the oddity flagged at iso.py:106-112)."""
    return np.concatenate([table, table[:1]])


# ----------------------------------------------------------------------------------------
# PCHIP as scipy builds it (site-packages/scipy/interpolate/_cubic.py:279-334): weighted
# harmonic mean slopes inside, shape-preserving three-point formula at both ends.
# ----------------------------------------------------------------------------------------
def _pchip_end_slope(h0: float, h1: float, m0: float, m1: float) -> float:
    d = ((2.0 * h0 + h1) * m0 - h0 * m1) / (h0 + h1)
    if np.sign(d) != np.sign(m0):
        return 0.0
    if np.sign(m0) != np.sign(m1) and abs(d) > 3.0 * abs(m0):
        return 3.0 * m0
    return float(d)


def pchip_slopes(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    h = np.diff(x)
    m = np.diff(y) / h
    n = len(x)
    d = np.zeros(n, dtype=np.float64)
    for k in range(1, n - 1):
        if m[k - 1] == 0.0 or m[k] == 0.0 or np.sign(m[k - 1]) != np.sign(m[k]):
            d[k] = 0.0
        else:
            w1 = 2.0 * h[k] + h[k - 1]
            w2 = h[k] + 2.0 * h[k - 1]
            d[k] = (w1 + w2) / (w1 / m[k - 1] + w2 / m[k])
    d[0] = _pchip_end_slope(h[0], h[1], m[0], m[1])
    d[-1] = _pchip_end_slope(h[-1], h[-2], m[-1], m[-2])
    return d


def pchip_eval(x: np.ndarray, y: np.ndarray, d: np.ndarray, q: np.ndarray) -> np.ndarray:
    """Cubic Hermite evaluation on the interval that contains each query point."""
    q = np.asarray(q, dtype=np.float64)
    i = np.clip(np.searchsorted(x, q, side="right") - 1, 0, len(x) - 2)
    h = x[i + 1] - x[i]
    s = q - x[i]
    m = (y[i + 1] - y[i]) / h
    # scipy stores the Hermite cubic in the power basis around x[i]
    c2 = (3.0 * m - 2.0 * d[i] - d[i + 1]) / h
    c3 = (d[i] + d[i + 1] - 2.0 * m) / (h * h)
    return y[i] + s * (d[i] + s * (c2 + s * c3))


def iso226_spl(phon: float, freqs_hz) -> np.ndarray:
    """Equal-loudness contour L_p(f; phon) in dB SPL (iso.py:86-173).

    Raises ValueError outside phon in [0,90] (iso.py:97-98) or f in [20,20000] (iso.py:152-153).
    The output keeps the *input dtype* (``np.zeros_like`` at iso.py:156), so integer
    frequencies give integer-truncated levels, exactly as the reference does.
    """
    if phon < 0 or phon > 90:
        raise ValueError("Phon must be in range [0, 90]")
    f_in = np.asarray(freqs_hz)
    if np.any(f_in < F_LO_HZ) or np.any(f_in > F_HI_HZ):
        raise ValueError("Frequency must be in [20, 20000] Hz")
    xk = _knots30()
    f = f_in.astype(np.float64)
    cols = []
    for tab in (ISO_ALPHA, ISO_LU, ISO_TF):
        yk = _wrapped(tab)
        cols.append(pchip_eval(xk, yk, pchip_slopes(xk, yk), f))
    alpha, lu, tf = cols
    a = 0.00447 * (10.0 ** (0.025 * phon) - 1.15)
    b = (0.4 * 10.0 ** ((tf + lu) / 10.0 - 9.0)) ** alpha
    spl = (10.0 / alpha) * np.log10(a + b) - lu + 94.0
    out = np.zeros_like(f_in)
    out[...] = spl                      # dtype-preserving store, see docstring
    return out


def iso226_weight_grid() -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(phon_knots[10], freq_knots[30], weights[10,30]) -- iso.py:176-199, :202-235."""
    fk = _knots30()
    spl = np.stack([iso226_spl(float(ph), fk) for ph in PHON_KNOTS])
    w = np.clip((1.0 - spl / spl.max()) ** 2, 0.0, 1.0)
    return PHON_KNOTS.copy(), fk, w


@dataclass
class WeightInterp:
    """Stand-in for the scipy RegularGridInterpolator the reference builds (iso.py:238-266):
    bilinear, ``bounds_error=False``, ``fill_value=1.0``; exposes ``grid``/``values`` like scipy."""
    grid: Tuple[np.ndarray, np.ndarray]
    values: np.ndarray
    fill_value: float = 1.0

    def __call__(self, pts) -> np.ndarray:
        pts = np.asarray(pts, dtype=np.float64)          # scipy casts queries to fp64
        out = np.empty(pts.shape[0], dtype=np.float64)
        (g0, g1), v = self.grid, self.values
        x0, x1 = pts[:, 0], pts[:, 1]
        oob = (x0 < g0[0]) | (x0 > g0[-1]) | (x1 < g1[0]) | (x1 > g1[-1])
        i0 = np.clip(np.searchsorted(g0, x0) - 1, 0, len(g0) - 2)
        i1 = np.clip(np.searchsorted(g1, x1) - 1, 0, len(g1) - 2)
        t0 = (x0 - g0[i0]) / (g0[i0 + 1] - g0[i0])
        t1 = (x1 - g1[i1]) / (g1[i1 + 1] - g1[i1])
        out[:] = (v[i0, i1] * (1 - t0) * (1 - t1) + v[i0, i1 + 1] * (1 - t0) * t1
                  + v[i0 + 1, i1] * t0 * (1 - t1) + v[i0 + 1, i1 + 1] * t0 * t1)
        out[oob] = self.fill_value
        return out


def build_weight_interpolator() -> WeightInterp:
    ph, fk, w = iso226_weight_grid()
    return WeightInterp((ph, fk), w)


def bin_frequencies(n_fft: int, sr: int) -> torch.Tensor:
    """rfft bin centres in fp32 the way ``torch.fft.rfftfreq(n, d=1/sr)`` makes them
    (projections.py:74,:97; build.py:331): arange(F) * fp32(1/(n*d))."""
    step = np.float32(1.0 / (n_fft * (1.0 / sr)))
    return torch.arange(n_fft // 2 + 1, dtype=torch.float32) * float(step)


def phon_threshold(n_fft: int, sr: int, max_phon_level: float) -> torch.Tensor:
    """Per-bin SPL threshold (1,F,1) fp32 -- build.py:325-348."""
    f = bin_frequencies(n_fft, sr).numpy()
    spl = iso226_spl(float(max_phon_level), np.clip(f, F_LO_HZ, F_HI_HZ))
    return torch.tensor(spl, dtype=torch.float32).view(1, -1, 1)


# ----------------------------------------------------------------------------------------
# STFT / ISTFT (src/core/fourier_transforms.py:4-41 -> torch.stft / torch.istft).
# ----------------------------------------------------------------------------------------
def hann_periodic(n: int) -> torch.Tensor:
    return torch.hann_window(n, periodic=True, dtype=torch.float32)


def stft(x: torch.Tensor, n_fft: int, hop: int) -> torch.Tensor:
    """centre=True (reflect pad n_fft/2), periodic Hann(n_fft), onesided, unnormalised.
    (B,T) -> (B, n_fft/2+1, 1+T//hop) complex64."""
    half = n_fft // 2
    xp = torch.nn.functional.pad(x.unsqueeze(1), (half, half), mode="reflect").squeeze(1)
    frames = xp.unfold(-1, n_fft, hop)                           # (B, T', n_fft)
    spec = torch.fft.rfft(frames * hann_periodic(n_fft).to(x.device), dim=-1)  # (B, T', F); window H2D as fourier_transforms.py:20
    return spec.transpose(1, 2)


def istft(spec: torch.Tensor, n_fft: int, hop: int) -> torch.Tensor:
    """Inverse of :func:`stft`: irfft, x window, overlap-add, / sum(window^2), trim n_fft/2.
    (B,F,T') -> (B, hop*(T'-1))."""
    B, _, n_frames = spec.shape
    w = hann_periodic(n_fft).to(spec.device)
    frames = torch.fft.irfft(spec.transpose(1, 2), n=n_fft, dim=-1) * w     # (B,T',n_fft)
    total = n_fft + hop * (n_frames - 1)
    fold = lambda cols: torch.nn.functional.fold(                    # noqa: E731
        cols, output_size=(1, total), kernel_size=(1, n_fft), stride=(1, hop)).reshape(-1, total)
    y = fold(frames.transpose(1, 2))
    env = fold((w * w).view(1, n_fft, 1).expand(1, n_fft, n_frames))
    half = n_fft // 2
    y, env = y[:, half:total - half], env[:, half:total - half]
    if bool(env.abs().min() < 1e-11):
        raise RuntimeError("window overlap add min: 1")
    return y / env


# ----------------------------------------------------------------------------------------
# Projections (src/core/projections.py).
# ----------------------------------------------------------------------------------------
def project_linf(p: torch.Tensor, eps: float) -> torch.Tensor:          # projections.py:37-39
    return p.clamp(-eps, eps)


def project_l2(p: torch.Tensor, eps: float) -> torch.Tensor:            # projections.py:41-46
    nrm = torch.linalg.vector_norm(p)
    if bool(nrm > eps):
        return p * (nrm.reciprocal() * eps)       # python ``eps / tensor`` is reciprocal()*eps
    return p


def project_snr(clean: torch.Tensor, p: torch.Tensor, snr_db: float) -> torch.Tensor:
    """projections.py:11-35.  NB the target norm uses ``clean.numel()`` even for a (1,T) p."""
    p_sig = (clean * clean).mean()
    p_noise = (p * p).mean()
    if bool(10.0 * torch.log10(p_sig / (p_noise + 1e-12)) >= snr_db):
        return p
    want = torch.sqrt(p_sig / (10.0 ** (snr_db / 10.0)) * clean.numel())
    have = torch.linalg.vector_norm(p.reshape(-1))
    if bool(have < 1e-8):
        return p
    return p * (want / have)


def total_variation(x: torch.Tensor) -> torch.Tensor:
    return (x[:, 1:] - x[:, :-1]).abs().sum()


def project_tv(p: torch.Tensor, clean: torch.Tensor, tv_epsilon: float) -> torch.Tensor:
    """projections.py:56-66."""
    budget = tv_epsilon * total_variation(clean)
    tv = total_variation(p)
    if bool(tv > budget):
        return p * (budget / tv)
    return p


def band_mask(n_fft: int, sr: int, f_min: float, f_max: float) -> torch.Tensor:
    """1 for bins strictly below f_min or strictly above f_max, else 0 (projections.py:74-77):
    the band [f_min, f_max] is *removed*."""
    f = bin_frequencies(n_fft, sr)
    return ((f < f_min) | (f > f_max)).to(torch.float32).view(1, -1, 1)


def fm_weighted_norm(spec: torch.Tensor, interp, n_fft: int, sr: int) -> torch.Tensor:
    """projections.py:83-113: sqrt(sum |X|^2 * w(10*log10(|X|^2+1e-10), f_k))."""
    B, F, Tn = spec.shape
    power = spec.abs() ** 2
    spl = 10.0 * torch.log10(power + 1e-10)
    f = bin_frequencies(n_fft, sr).to(spec.device).view(1, F, 1).expand(B, F, Tn)
    q = torch.stack([spl, f], dim=-1).reshape(-1, 2).detach().cpu().numpy()      # the D2H of projections.py:104
    w = torch.tensor(interp(q).reshape(B, F, Tn), dtype=torch.float32, device=spec.device)
    return torch.sqrt((power * w).sum())


def project_fm(spec: torch.Tensor, interp, n_fft: int, sr: int, fm_epsilon: float) -> torch.Tensor:
    """projections.py:116-133."""
    nrm = fm_weighted_norm(spec, interp, n_fft, sr)
    if bool(nrm <= fm_epsilon):
        return spec
    return spec * (nrm.clamp(min=1e-8).reciprocal() * fm_epsilon)


def phon_levels(spec: torch.Tensor) -> torch.Tensor:
    """Per-bin level in dB: 20*log10(|X|+1e-8) (projections.py:142-143)."""
    return 20.0 * torch.log10(spec.abs() + 1e-8)


def scaled_threshold(spl_thresh: torch.Tensor, phon_reference_db: float) -> torch.Tensor:
    return spl_thresh - spl_thresh.max() + phon_reference_db             # projections.py:147


def project_phon(spec: torch.Tensor, spl_thresh: torch.Tensor, phon_reference_db: float) -> torch.Tensor:
    """projections.py:138-159.  Every bin goes through the dB round trip, clipped or not."""
    level = phon_levels(spec)
    thr = scaled_threshold(spl_thresh, phon_reference_db)
    level = torch.where(level > thr, thr, level)
    mag = 10 ** (level / 20)
    return mag * torch.exp(1j * spec.angle())


# ----------------------------------------------------------------------------------------
# Dispatch (src/training_utils/train.py:27-99) and the two optimiser steps (train.py:155-175).
# ----------------------------------------------------------------------------------------
FREQ_NORMS = ("fletcher_munson", "min_max_freqs", "max_phon")
TIME_NORMS = ("l2", "linf", "snr", "tv")


@dataclass
class Hyper:
    """The hot-path fields of the reference's argparse Namespace (parser.py:10-61 defaults)."""
    norm_type: str = "max_phon"
    lr: float = 1e-4
    optimizer_type: str = "adam"
    attack_mode: str = "untargeted"
    fm_epsilon: float = 2.0
    l2_size: float = 0.05
    linf_size: float = 1e-4
    snr_db: float = 64.0
    min_freq_attack: float = 120.0
    max_freq_attack: float = 20000.0
    tv_epsilon: float = 1e-3
    max_phon_level: float = 20.0
    phon_reference_db: float = 65.0
    sr: int = 16000
    n_fft: int = 1024
    hop_length: int = 256
    win_length: int = 1024
    target: str = "delete"
    target_reps: int = 5
    device: str = "cpu"


def match_length(x: torch.Tensor, length: int) -> torch.Tensor:         # train.py:27-35
    have = x.shape[-1]
    if have < length:
        return torch.nn.functional.pad(x, (0, length - have))
    return x[..., :length]


def constrain(p: torch.Tensor, clean: Optional[torch.Tensor], hp, interp=None,
              spl_thresh: Optional[torch.Tensor] = None) -> torch.Tensor:
    """train.py:69-99 (+ :38-66 for the three STFT-domain norms)."""
    kind = hp.norm_type
    with torch.no_grad():
        if kind in FREQ_NORMS:
            spec = stft(p, hp.n_fft, hp.hop_length)
            if kind == "min_max_freqs":
                spec = spec * band_mask(hp.n_fft, hp.sr, hp.min_freq_attack, hp.max_freq_attack).to(spec.device)
            elif kind == "fletcher_munson":
                spec = project_fm(spec, interp, hp.n_fft, hp.sr, hp.fm_epsilon)
            else:
                spec = project_phon(spec, spl_thresh, hp.phon_reference_db)
            out = istft(spec, hp.n_fft, hp.hop_length)
            return out if clean is None else match_length(out, clean.shape[-1])
        if kind == "l2":
            return project_l2(p, hp.l2_size)
        if kind == "linf":
            return project_linf(p, hp.linf_size)
        if kind == "snr":
            if clean is None:
                raise ValueError("snr projection needs clean_audio")
            return project_snr(clean, p, hp.snr_db)
        if kind == "tv":
            if clean is None:
                raise ValueError("tv projection needs clean_audio")
            return project_tv(p, clean, hp.tv_epsilon)
    raise ValueError(f"Unknown norm_type: {kind!r}")


def pgd_step(p: torch.Tensor, grad: torch.Tensor, lr: float) -> torch.Tensor:
    """train.py:161 -- ``p.add_(lr * p.grad.sign())`` (sign(0)=0, sign(nan)=0)."""
    return p + lr * grad.sign()


@dataclass
class AdamState:
    """torch.optim.Adam defaults as build.py:357 creates it (betas .9/.999, eps 1e-8, no decay)."""
    m: torch.Tensor
    v: torch.Tensor
    t: int = 0
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8


def adam_step(p: torch.Tensor, grad: torch.Tensor, st: AdamState, lr: float) -> torch.Tensor:
    """site-packages/torch/optim/adam.py single-tensor arithmetic (:457-547); mutates ``st``."""
    st.t += 1
    st.m.lerp_(grad, 1.0 - st.beta1)
    st.v.mul_(st.beta2).addcmul_(grad, grad, value=1.0 - st.beta2)
    bc1 = 1.0 - st.beta1 ** st.t
    bc2 = 1.0 - st.beta2 ** st.t
    denom = (st.v.sqrt() / (bc2 ** 0.5)).add_(st.eps)
    return torch.addcdiv(p, st.m, denom, value=-(lr / bc1))


def step_and_constrain(p, grad, clean, hp, interp=None, spl_thresh=None,
                       adam: Optional[AdamState] = None, lr: Optional[float] = None) -> torch.Tensor:
    """One hot-path iteration given the gradient the model produced (train.py:155-175)."""
    lr = hp.lr if lr is None else lr
    if hp.optimizer_type == "pgd":
        q = pgd_step(p, grad, lr)
    elif hp.optimizer_type == "adam":
        q = adam_step(p, grad, adam, lr)
    else:
        raise NotImplementedError(f"Optimization type not implemented: {hp.optimizer_type!r}")
    return constrain(q, clean, hp, interp, spl_thresh)


# ----------------------------------------------------------------------------------------
# Word error counters (what HF ``evaluate``'s "wer" on jiwer returns: sum(S+D+I)/sum(ref words)).
# ----------------------------------------------------------------------------------------
def edit_counts(ref: Sequence[str], hyp: Sequence[str]) -> Tuple[int, int]:
    errors = words = 0
    for r, h in zip(ref, hyp):
        a, b = r.split(), h.split()
        prev = list(range(len(b) + 1))
        for i, wa in enumerate(a, 1):
            cur = [i] + [0] * len(b)
            for j, wb in enumerate(b, 1):
                cur[j] = min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (wa != wb))
            prev = cur
        errors += prev[-1]
        words += len(a)
    return errors, words


class WerMetric:
    """``.compute(predictions=, references=)`` like the object run_attack.py:27 loads."""

    def compute(self, predictions, references) -> float:
        e, w = edit_counts(references, predictions)
        return e / max(w, 1)


# ----------------------------------------------------------------------------------------
# The attack step around the hot path (train.py:126-175), used as CPU baseline / reference arm.
# ----------------------------------------------------------------------------------------
VOCAB = ["<pad>", "<s>", "</s>", "<unk>", "|"] + list("ETAONIHSRDLUMWCFGYPBVK'XJQZ")


def text_to_labels(texts: Sequence[str]) -> torch.Tensor:
    """What ``processor(text=..., padding=True).input_ids`` yields for the 32-token wav2vec2
    vocabulary, with pads replaced by -100 (loss_helpers.py:7-9,:17-20)."""
    import re
    table = {c: i for i, c in enumerate(VOCAB)}
    rows = []
    for t in texts:
        t = re.sub(r"\s+", " ", t.replace("<unk>", "").lower()).strip().upper()
        rows.append([table.get("|" if c == " " else c, table["<unk>"]) for c in t])
    width = max(len(r) for r in rows)
    return torch.tensor([r + [-100] * (width - len(r)) for r in rows], dtype=torch.long)


def greedy_transcripts(logits: torch.Tensor) -> list:
    """argmax -> CTC collapse -> text (loss_helpers.py:25-28), lower-cased and stripped."""
    out = []
    for row in logits.argmax(-1).tolist():
        chars, last = [], None
        for tok in row:
            if tok != last and tok > 3:
                chars.append(" " if tok == 4 else VOCAB[tok])
            last = tok
        out.append(" ".join("".join(chars).split()).lower())
    return out


def attack_iteration(model, p, clean, texts, hp, interp=None, spl_thresh=None, adam=None, lr=None):
    """compose -> wav2vec2+CTC -> backward -> step -> projection.  Returns (p_new, loss, transcripts)."""
    direction = 1.0 if hp.attack_mode == "untargeted" else -1.0
    if hp.attack_mode == "targeted":
        texts = [" ".join([hp.target] * hp.target_reps)] * clean.shape[0]
    p = p.detach().clone().requires_grad_(True)
    x_adv = (clean + p).clamp_(-1.0, 1.0)
    out = model(input_values=x_adv, labels=text_to_labels(texts).to(clean.device))
    sign = direction if hp.optimizer_type == "pgd" else -direction
    (sign * out.loss).backward()
    with torch.no_grad():
        p_new = step_and_constrain(p.detach(), p.grad, clean, hp, interp, spl_thresh, adam, lr)
    return p_new.detach(), float(out.loss.detach()), greedy_transcripts(out.logits.detach())
