"""ctypes binding of libpaa.so (include/paa.h) and the little runtime PyTorch needs around it:
one handle per (device, n_fft, hop, sr), a scratch buffer per handle, the current CUDA stream.

PyTorch is plumbing here -- device memory, streams -- the arithmetic is in the library.
There is deliberately no fallback: a missing library or a CPU tensor raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Dict, Optional, Tuple

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# PAA_LIBPAA: another build of the same library (kernel experiments under tools/); the default is the in-tree one
LIB_PATH = os.environ.get("PAA_LIBPAA") or os.path.join(_HERE, "libpaa.so")

# status codes of include/paa.h
OK, ERR_NULL, ERR_SHAPE, ERR_UNSUPPORTED, ERR_CUDA, ERR_RANGE, ERR_NEED_CLEAN, ERR_ALIAS, ERR_NOLA, ERR_STATE = range(10)
STEP_NONE, STEP_PGD, STEP_ADAM = 0, 1, 2
S_SCALE, S_NORM, S_AUX0, S_AUX1 = 0, 1, 2, 3

EXPORTS = (
    "paa_status_string paa_version paa_launch_count paa_create paa_destroy paa_last_cuda_error paa_num_bins paa_num_frames "
    "paa_scratch_bytes paa_scalars paa_iso226_spl paa_weight_grid paa_spl_thresh paa_interp2 paa_set_fm_grid "
    "paa_project_linf paa_project_l2 paa_project_snr paa_project_tv paa_project_min_max_freqs "
    "paa_project_max_phon paa_project_fletcher_munson paa_step_only paa_stft paa_istft "
    "paa_spec_min_max_freqs paa_spec_phon_level paa_spec_fm_norm paa_spec_fm_project paa_compose_clamp "
    "paa_compose_clamp_backward paa_clean_stats "
    "paa_wer_counts"
).split()


MAX_PARTS = 8


class Parts(C.Structure):
    """struct paa_parts -- mode U: per-rank partial gradients / clean statistics in peer-mapped device memory"""
    _fields_ = [("n", C.c_int), ("grad", C.c_void_p * MAX_PARTS), ("clean_stats", C.c_void_p * MAX_PARTS),
                ("clean_numel", C.c_int64)]


class Step(C.Structure):
    """struct paa_step"""
    _fields_ = [("mode", C.c_int), ("grad", C.c_void_p), ("lr", C.c_double), ("adam_m", C.c_void_p),
                ("adam_v", C.c_void_p), ("adam_t", C.c_int64), ("beta1", C.c_double), ("beta2", C.c_double),
                ("eps", C.c_double), ("parts", C.POINTER(Parts))]


def _load() -> C.CDLL:
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). paa_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_double
    dp = C.POINTER(C.c_double)
    sp = C.POINTER(Step)
    sig = {
        "paa_status_string": (C.c_char_p, [i32]),
        "paa_version": (i32, []),
        "paa_launch_count": (i64, []),
        "paa_create": (i32, [i32, i32, i32, i32, C.POINTER(vp)]),
        "paa_destroy": (i32, [vp]),
        "paa_last_cuda_error": (i32, [vp]),
        "paa_num_bins": (i32, [vp]),
        "paa_num_frames": (i32, [vp, i32]),
        "paa_scratch_bytes": (C.c_size_t, [vp, i32, i32]),
        "paa_scalars": (i32, [vp, vp, C.POINTER(C.c_float), vp]),
        "paa_iso226_spl": (i32, [f64, dp, i32, dp]),
        "paa_weight_grid": (i32, [dp, dp, dp]),
        "paa_spl_thresh": (i32, [i32, i32, f64, C.POINTER(C.c_float)]),
        "paa_interp2": (i32, [dp, i32, dp, i32, dp, f64, dp, i32, dp]),
        "paa_set_fm_grid": (i32, [vp, dp, i32, dp, i32, dp, f64]),
        "paa_project_linf": (i32, [vp, vp, vp, i32, i32, f64, f64, sp, vp]),
        "paa_project_l2": (i32, [vp, vp, vp, i32, i32, f64, sp, vp, vp]),
        "paa_project_snr": (i32, [vp, vp, vp, i32, i32, vp, i64, f64, sp, vp, vp]),
        "paa_project_tv": (i32, [vp, vp, vp, i32, i32, vp, i32, i32, f64, sp, vp, vp]),
        "paa_project_min_max_freqs": (i32, [vp, vp, vp, i32, i32, i32, f64, f64, sp, vp, vp]),
        "paa_project_max_phon": (i32, [vp, vp, vp, i32, i32, i32, vp, f64, sp, vp, vp]),
        "paa_project_fletcher_munson": (i32, [vp, vp, vp, i32, i32, i32, f64, i32, sp, vp, vp]),
        "paa_step_only": (i32, [vp, vp, vp, i32, i32, sp, vp]),
        "paa_stft": (i32, [vp, vp, i32, i32, vp, i64, i64, i64, vp]),
        "paa_istft": (i32, [vp, vp, i64, i64, i64, i32, i32, vp, vp]),
        "paa_spec_min_max_freqs": (i32, [vp, vp, vp, i32, i32, i64, i64, i64, f64, f64, vp]),
        "paa_spec_phon_level": (i32, [vp, vp, vp, i32, i32, i64, i64, i64, vp, f64, vp]),
        "paa_spec_fm_norm": (i32, [vp, vp, i32, i32, i64, i64, i64, vp, vp]),
        "paa_spec_fm_project": (i32, [vp, vp, vp, i32, i32, i64, i64, i64, f64, vp, vp]),
        "paa_compose_clamp": (i32, [vp, vp, i32, vp, i32, i32, vp, vp]),
        "paa_compose_clamp_backward": (i32, [vp, vp, i32, vp, i32, i32, vp, vp, vp]),
        "paa_clean_stats": (i32, [vp, vp, i32, i32, vp, vp, vp]),
        "paa_wer_counts": (i32, [C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), i32, C.POINTER(i64), C.POINTER(i64)]),
    }
    for name in EXPORTS:
        fn = getattr(lib, name)          # AttributeError here = a symbol of paa.h is not exported
        fn.restype, fn.argtypes = sig[name]
    return lib


lib = _load()


class PaaError(RuntimeError):
    pass


def check(status: int, handle=None) -> None:
    """Map a paa_status onto the exception type the reference raises in the same situation."""
    if status == OK:
        return
    msg = lib.paa_status_string(status).decode()
    if status == ERR_NEED_CLEAN:
        raise ValueError(msg)                                   # train.py:90-95
    if status == ERR_RANGE:
        raise ValueError(msg)                                   # iso.py:97-98, :152-153
    if status == ERR_NOLA:
        raise RuntimeError(msg)                                 # torch.istft's own check
    if status == ERR_CUDA and handle is not None:
        msg += f" (cudaError {lib.paa_last_cuda_error(handle)})"
    raise PaaError(f"libpaa: {msg} [status {status}]")


# ---- host tables ------------------------------------------------------------------------------
def _dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def iso226_spl(phon: float, freqs) -> np.ndarray:
    f = np.ascontiguousarray(np.asarray(freqs, dtype=np.float64).reshape(-1))
    out = np.empty_like(f)
    check(lib.paa_iso226_spl(float(phon), _dptr(f), f.size, _dptr(out)))
    return out.reshape(np.shape(freqs))


def weight_grid() -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    ph, fk, w = np.empty(10), np.empty(30), np.empty((10, 30))
    check(lib.paa_weight_grid(_dptr(ph), _dptr(fk), _dptr(w)))
    return ph, fk, w


def spl_thresh(n_fft: int, sr: int, max_phon_level: float) -> np.ndarray:
    out = np.empty(n_fft // 2 + 1, dtype=np.float32)
    check(lib.paa_spl_thresh(int(n_fft), int(sr), float(max_phon_level), out.ctypes.data_as(C.POINTER(C.c_float))))
    return out


def interp2(g0, g1, values, fill, query) -> np.ndarray:
    g0, g1 = (np.ascontiguousarray(x, dtype=np.float64) for x in (g0, g1))
    v = np.ascontiguousarray(values, dtype=np.float64)
    q = np.ascontiguousarray(np.asarray(query, dtype=np.float64).reshape(-1, 2))
    out = np.empty(q.shape[0])
    check(lib.paa_interp2(_dptr(g0), g0.size, _dptr(g1), g1.size, _dptr(v), float(fill), _dptr(q), q.shape[0], _dptr(out)))
    return out


def wer_counts(references, hypotheses) -> Tuple[int, int]:
    n = len(references)
    if len(hypotheses) != n:
        raise ValueError("references and hypotheses differ in length")
    arr = C.c_char_p * n
    r = arr(*[s.encode() for s in references])
    h = arr(*[s.encode() for s in hypotheses])
    e, w = C.c_int64(0), C.c_int64(0)
    check(lib.paa_wer_counts(r, h, n, C.byref(e), C.byref(w)))
    return int(e.value), int(w.value)


# ---- device runtime ---------------------------------------------------------------------------
class Plan:
    """A libpaa handle plus its caller-owned scratch, for one (device, n_fft, hop, sr)."""

    def __init__(self, device: torch.device, n_fft: int, hop: int, sr: int):
        self.device, self.n_fft, self.hop, self.sr = device, int(n_fft), int(hop), int(sr)
        h = C.c_void_p()
        check(lib.paa_create(device.index, self.n_fft, self.hop, self.sr, C.byref(h)))
        self.h = h
        # caller-owned scratch of the C ABI, one buffer per CUDA stream that has used this plan: concurrent calls on
        # different streams never share partials / scalars / staging (include/paa.h: re-entrant with distinct scratch)
        self._scratch_by_stream = {}
        self._fm_obj = None              # the installed interpolator itself: a live reference, so its id cannot be reused
        self._fm_sig = None

    def __del__(self):
        try:
            lib.paa_destroy(self.h)
        except Exception:
            pass

    def _scratch_tensor(self, need: int = 0) -> torch.Tensor:
        key = stream_ptr(self.device)
        buf = self._scratch_by_stream.get(key)
        if buf is None or buf.numel() < need:
            with torch.cuda.device(self.device):
                buf = torch.empty(max(need, 256), dtype=torch.uint8, device=self.device)   # allocated on the current stream
            self._scratch_by_stream.pop(key, None)
            self._scratch_by_stream[key] = buf
            while len(self._scratch_by_stream) > 4:          # ephemeral streams: keep the four most recent buffers (a
                self._scratch_by_stream.pop(next(iter(self._scratch_by_stream)))     # dropped one is freed in stream order)
        return buf

    @property
    def _scratch(self) -> torch.Tensor:
        """The scratch buffer of the current stream (scalars of its last reducing call live at the front)."""
        return self._scratch_tensor()

    def scratch(self, rows: int, T: int) -> int:
        need = int(lib.paa_scratch_bytes(self.h, int(rows), int(T)))
        return self._scratch_tensor(need).data_ptr()

    def scalars(self) -> np.ndarray:
        """The PAA_S_* scalars of the last reducing call (synchronises the current stream)."""
        out = (C.c_float * 8)()
        check(lib.paa_scalars(self.h, self._scratch.data_ptr(), out, stream_ptr(self.device)), self.h)
        return np.array(out[:], dtype=np.float32)

    def set_fm_grid(self, interp) -> None:
        """Install interp.grid / interp.values (the object iso.py:238-266 returns).  The same object is installed
        once; a different object is compared by content (300 doubles) and re-installed only if it differs."""
        if interp is self._fm_obj:
            return
        g0, g1 = (np.ascontiguousarray(g, dtype=np.float64) for g in interp.grid)
        v = np.ascontiguousarray(interp.values, dtype=np.float64)
        fill = 1.0 if getattr(interp, "fill_value", 1.0) is None else float(interp.fill_value)
        sig = (g0.tobytes(), g1.tobytes(), v.tobytes(), fill)
        if sig != self._fm_sig:
            check(lib.paa_set_fm_grid(self.h, _dptr(g0), g0.size, _dptr(g1), g1.size, _dptr(v), fill), self.h)
            self._fm_sig = sig
        self._fm_obj = interp


_plans: Dict[Tuple[int, int, int, int], Plan] = {}
_lock = threading.Lock()


def plan_for(t: torch.Tensor, args) -> Plan:
    need_cuda(t)
    n_fft, hop, sr = int(args.n_fft), int(args.hop_length), int(args.sr)
    win = int(getattr(args, "win_length", n_fft))
    if win != n_fft:
        # the reference builds its window with n_fft samples; torch.stft rejects a mismatch (fourier_transforms.py:20-27)
        raise RuntimeError(f"win_length ({win}) must equal n_fft ({n_fft})")
    key = (t.device.index, n_fft, hop, sr)
    with _lock:
        p = _plans.get(key)
        if p is None:
            p = _plans[key] = Plan(t.device, n_fft, hop, sr)
    return p


def plan_plain(t: torch.Tensor) -> Plan:
    """Time-domain projections do not depend on the STFT geometry; any plan of the device will do."""
    need_cuda(t)
    with _lock:
        for k, p in _plans.items():
            if k[0] == t.device.index:
                return p
        p = _plans[(t.device.index, 1024, 256, 16000)] = Plan(t.device, 1024, 256, 16000)
    return p


def need_cuda(*tensors) -> None:
    for t in tensors:
        if t is None:
            continue
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise RuntimeError("paa_b200 runs on CUDA tensors only (no CPU fallback); got "
                               f"{type(t).__name__} on {getattr(t, 'device', '?')}")


def f32c(t: torch.Tensor) -> torch.Tensor:
    """fp32, contiguous view of `t` (copy only if it has to)."""
    if t.dtype != torch.float32:
        raise TypeError(f"expected float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def make_parts(grad_ptrs=None, stat_ptrs=None, clean_numel: int = 0) -> Parts:
    """paa_parts from raw device pointers (one per rank, rank order)."""
    n = len(grad_ptrs or stat_ptrs or ())
    if not 0 < n <= MAX_PARTS:
        raise ValueError(f"mode U supports 1..{MAX_PARTS} ranks, got {n}")
    p = Parts()
    p.n = n
    for k in range(n):
        p.grad[k] = int(grad_ptrs[k]) if grad_ptrs else None
        p.clean_stats[k] = int(stat_ptrs[k]) if stat_ptrs else None
    p.clean_numel = int(clean_numel)
    return p


def make_step(mode: int = STEP_NONE, grad: Optional[torch.Tensor] = None, lr: float = 0.0,
              m: Optional[torch.Tensor] = None, v: Optional[torch.Tensor] = None, t: int = 0,
              betas=(0.9, 0.999), eps: float = 1e-8, parts: Optional[Parts] = None):
    if mode == STEP_NONE and parts is None:
        return None
    st = Step(mode, grad.data_ptr() if grad is not None else None, float(lr), m.data_ptr() if m is not None else None,
              v.data_ptr() if v is not None else None, int(t), float(betas[0]), float(betas[1]), float(eps),
              C.pointer(parts) if parts is not None else None)
    st._keep = parts                       # the pointer does not own the struct
    return st


def attach_parts(step: Step, parts: Optional[Parts]) -> None:
    if parts is not None:
        step.parts = C.pointer(parts)
        step._keep = parts


def step_ref(step):
    return C.byref(step) if step is not None else None
