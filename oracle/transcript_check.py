"""TEST INFRASTRUCTURE (like everything under oracle/): transcript / WER identity of an attack run whose step +
projection is libpaa against the same run with the reference's torch arithmetic (oracle/paa_oracle.py executed on CUDA
tensors), on the same GPU and the same model.  Used by tests/test_gpu_transcripts.py and by bench.py's baseline leg
(`transcript_identity` in the bench line); never on the product path.

north_star: "an identical post-attack transcript and WER"  (reference: src/core/loss_helpers.py:25-32 argmax ->
batch_decode -> WER, src/training_utils/train.py:149-153).

Two comparisons per step k:

  teacher-forced  both implementations get the SAME perturbation p_k and the SAME gradient g_k (the reference
                  trajectory's); their outputs p_{k+1} differ by the projection's fp32 rounding (<= 1e-5 relative), and
                  the greedy token of every logit frame of model(clean + p_{k+1}) is compared.  This is "identical inputs
                  and seeds" in the strict sense: any flip is caused by the hot path alone.
  free-running    each implementation follows its own trajectory for all steps (what a user sees after N steps).  Here a
                  rounding-level difference in p can flip sign(g) where |g| ~ 0, so the two perturbations drift apart by
                  2*lr on a small fraction of samples; the transcripts are compared all the same.

A *flip* is a logit frame whose argmax differs; its *margin* is the reference's top-1 minus top-2 logit on that frame
(how decisive the frame was).  A random-init wav2vec2 has nearly flat logits (std ~0.5, top-2 margins down to 1e-7), so
near-ties exist; the report carries every flip with the smallest and largest margin seen, and the caller asserts "no
flip above margin M".

What sets M is the gradient source, not the hot path: torch runs cuDNN convolutions in TF32 by default
(torch.backends.cudnn.allow_tf32 = True, the setting the reference inherits), and TF32 rounds the convolution inputs to
10 mantissa bits -- a last-bit fp32 difference in one waveform sample that crosses a TF32 rounding boundary becomes a
5e-4 RELATIVE change of that sample, and the logits move by ~1e-3.  Two controls make that visible:

  control_one_ulp   the reference's own p_{k+1} with a random half of its samples moved to the adjacent fp32 value (what
                    any other correct fp32 implementation -- another torch version, another reduction order -- could
                    produce) against the unmodified one: same flip statistics as libpaa's.
  cudnn_tf32=False  the same comparison with the gradient source in full fp32: the logit differences fall to the 1e-6
                    range and the flips with them.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import paa_oracle as orc


def _forward(model, clean, p, labels, micro, want_grad, sign):
    """model(clamp(clean + p)) in chunks of `micro` utterances; returns (loss, logits, grad or None)."""
    B = clean.shape[0]
    micro = B if not micro else micro
    per_row = p.shape[0] == B and B > 1
    q = p.detach().clone().requires_grad_(want_grad)
    total, chunks = 0.0, []
    for lo in range(0, B, micro):
        hi = min(lo + micro, B)
        x = (clean[lo:hi] + (q[lo:hi] if per_row else q)).clamp_(-1.0, 1.0)
        with torch.set_grad_enabled(want_grad):
            out = model(input_values=x, labels=labels[lo:hi])
        if want_grad:
            (sign * out.loss).backward()
        total += float(out.loss.detach())
        chunks.append(out.logits.detach())
    return total, torch.cat(chunks, 0), (q.grad if want_grad else None)


def _compare(lg_ref, lg_new, acc, texts):
    ids_r, ids_n = lg_ref.argmax(-1), lg_new.argmax(-1)
    diff = ids_r != ids_n
    acc["frames"] += ids_r.numel()
    acc["max_logit_diff"] = max(acc["max_logit_diff"], float((lg_ref - lg_new).abs().max()))
    n = int(diff.sum())
    if n:
        top2 = lg_ref.topk(2, dim=-1).values
        margin = (top2[..., 0] - top2[..., 1])[diff]
        acc["flips"] += n
        acc["max_flip_margin"] = max(acc["max_flip_margin"] or 0.0, float(margin.max()))
        acc["min_flip_margin"] = min(acc["min_flip_margin"] if acc["min_flip_margin"] is not None else 1e30, float(margin.min()))
    hyp_r, hyp_n = orc.greedy_transcripts(lg_ref), orc.greedy_transcripts(lg_new)
    refs = [t.lower() for t in texts]
    same_text = hyp_r == hyp_n
    acc["transcript_mismatch_steps"] += 0 if same_text else 1
    import paa_b200.paa_lib as L
    wer_new = L.wer_counts(refs, hyp_n)                    # libpaa's C++ counters on libpaa's transcripts
    wer_ref = orc.edit_counts(refs, hyp_r)                 # the oracle's DP on the reference transcripts
    acc["wer_mismatch_steps"] += 0 if tuple(wer_new) == tuple(wer_ref) else 1
    acc["wer_last"] = {"errors": int(wer_new[0]), "ref_words": int(wer_new[1])}


def _blank():
    return {"frames": 0, "flips": 0, "min_flip_margin": None, "max_flip_margin": None, "max_logit_diff": 0.0,
            "transcript_mismatch_steps": 0, "wer_mismatch_steps": 0, "wer_last": None}


def one_ulp_nudge(p: torch.Tensor, seed: int) -> torch.Tensor:
    """A random half of the samples moved to the adjacent fp32 value (direction random)."""
    g = torch.Generator(device=p.device).manual_seed(seed)
    r = torch.rand(p.shape, generator=g, device=p.device)
    up = torch.nextafter(p, torch.full_like(p, float("inf")))
    dn = torch.nextafter(p, torch.full_like(p, float("-inf")))
    return torch.where(r < 0.25, up, torch.where(r < 0.5, dn, p))


def run(model, clean: torch.Tensor, texts, args, hp, steps: int, p0: torch.Tensor, interp_cpu=None, interp_gpu=None,
        spl_thresh=None, micro: int = 0, free_running: bool = True, control: bool = True,
        cudnn_tf32: Optional[bool] = None) -> Dict:
    """`args` drives libpaa (paa_b200.step_and_project), `hp` the oracle; both describe the same attack.
    PGD and Adam are supported (Adam keeps one state per trajectory).  `cudnn_tf32`: None = torch's default (True)."""
    prev_tf32 = torch.backends.cudnn.allow_tf32
    if cudnn_tf32 is not None:
        torch.backends.cudnn.allow_tf32 = bool(cudnn_tf32)
    try:
        return _run(model, clean, texts, args, hp, steps, p0, interp_cpu, interp_gpu, spl_thresh, micro, free_running, control)
    finally:
        torch.backends.cudnn.allow_tf32 = prev_tf32


def _run(model, clean, texts, args, hp, steps, p0, interp_cpu, interp_gpu, spl_thresh, micro, free_running, control) -> Dict:
    import paa_b200
    from paa_b200.training_utils import build as pbuild

    dev = clean.device
    direction = 1.0 if hp.attack_mode == "untargeted" else -1.0
    label_texts = [" ".join([hp.target] * hp.target_reps)] * clean.shape[0] if hp.attack_mode == "targeted" else list(texts)
    labels = orc.text_to_labels(label_texts).to(dev)
    sign = direction if hp.optimizer_type == "pgd" else -direction
    adam = hp.optimizer_type == "adam"

    p_ref = orc.constrain(p0, clean, hp, interp_cpu, spl_thresh)
    p_free = paa_b200.perturbation_constraint(p0, clean, args, interp_gpu, spl_thresh)
    st_ref = orc.AdamState(torch.zeros_like(p0), torch.zeros_like(p0)) if adam else None
    # teacher-forced libpaa gets the reference trajectory's optimiser state: a copy refreshed every step
    pa_tf = torch.nn.Parameter(p_ref.clone()) if adam else None
    opt_tf = pbuild.create_optimizer(args, pa_tf)[0] if adam else None
    pa_free = torch.nn.Parameter(p_free.clone()) if adam else None
    opt_free = pbuild.create_optimizer(args, pa_free)[0] if adam else None

    tf, fr, ctl = _blank(), _blank(), _blank()
    max_rel_p = 0.0
    same_input_diff = 0.0
    for k in range(steps):
        _, _, g_ref = _forward(model, clean, p_ref, labels, micro, True, sign)
        with torch.no_grad():
            if adam:      # hand libpaa the reference trajectory's state (m, v, t) before the step
                s = opt_tf._state_for(pa_tf)
                s["exp_avg"].copy_(st_ref.m); s["exp_avg_sq"].copy_(st_ref.v); s["step"].fill_(float(st_ref.t))
                pa_tf.data.copy_(p_ref)
                p_tf = paa_b200.step_and_project(pa_tf.data, g_ref, clean, args, interp_gpu, spl_thresh, optimizer=opt_tf)
            else:
                p_tf = paa_b200.step_and_project(p_ref, g_ref, clean, args, interp_gpu, spl_thresh)
            p_ref = orc.step_and_constrain(p_ref, g_ref, clean, hp, interp_cpu, spl_thresh, adam=st_ref)
        max_rel_p = max(max_rel_p, float((p_tf - p_ref).abs().max() / p_ref.abs().max()))
        _, lg_ref, _ = _forward(model, clean, p_ref, labels, micro, False, sign)
        _, lg_tf, _ = _forward(model, clean, p_tf, labels, micro, False, sign)
        _compare(lg_ref, lg_tf, tf, texts if hp.attack_mode == "untargeted" else label_texts)
        if control:
            _, lg_ctl, _ = _forward(model, clean, one_ulp_nudge(p_ref, 1000 + k), labels, micro, False, sign)
            _compare(lg_ref, lg_ctl, ctl, texts if hp.attack_mode == "untargeted" else label_texts)
            if k == 0:          # is the forward pass itself deterministic on identical input?
                _, lg_again, _ = _forward(model, clean, p_ref, labels, micro, False, sign)
                same_input_diff = float((lg_again - lg_ref).abs().max())
        if free_running:
            cur = pa_free.data if adam else p_free
            _, _, g_free = _forward(model, clean, cur, labels, micro, True, sign)
            with torch.no_grad():
                nxt = paa_b200.step_and_project(cur, g_free, clean, args, interp_gpu, spl_thresh, optimizer=opt_free)
            if adam:
                pa_free.data = nxt
            else:
                p_free = nxt
            _, lg_free, _ = _forward(model, clean, nxt, labels, micro, False, sign)
            _compare(lg_ref, lg_free, fr, texts if hp.attack_mode == "untargeted" else label_texts)
    out = {"steps": steps, "batch": int(clean.shape[0]), "p_rows": int(p0.shape[0]), "norm_type": hp.norm_type,
           "optimizer": hp.optimizer_type, "cudnn_tf32": bool(torch.backends.cudnn.allow_tf32),
           "max_rel_err_p_teacher_forced": float(f"{max_rel_p:.3e}"), "teacher_forced": tf}
    if control:
        ctl["same_input_logit_diff"] = same_input_diff
        out["control_one_ulp"] = ctl
    if free_running:
        p_last = pa_free.data if adam else p_free
        fr["rel_l2_p_vs_reference"] = float(f"{float((p_last - p_ref).norm() / p_ref.norm()):.3e}")
        out["free_running"] = fr
    return out


def flips_above(report: Dict, margin: float) -> bool:
    """True when some frame whose reference margin exceeds `margin` changed its token."""
    m = report["max_flip_margin"]
    return m is not None and m > margin
