"""The per-iteration perturbation update with the reference's entry points
(src/training_utils/train.py:27-182): ``perturbation_constraint`` keeps its signature, and
``step_and_project`` is the fused form (optimiser step + projection in the same kernels) that
``train_epoch`` uses."""
import logging
import time
from dataclasses import dataclass
from datetime import datetime
from typing import Iterable, Optional

import torch

try:
    from .. import paa_lib as L
    from ..core import compose, projections
except ImportError:
    import paa_lib as L
    from core import compose, projections

logger = logging.getLogger("asr_attack")

FREQ_NORMS = ("fletcher_munson", "min_max_freqs", "max_phon")


@dataclass(frozen=True)
class TrainEpochResult:
    p: torch.Tensor
    avg_ctc: float
    avg_wer: float

    def __iter__(self):          # lets ``p, ctc, wer = train_epoch(...)`` work (run_attack.py:64)
        return iter((self.p, self.avg_ctc, self.avg_wer))


def _avg(values: Iterable[float]) -> float:
    vals = list(values)
    return sum(vals) / max(len(vals), 1)


def _frequency_domain(p, clean_audio, args, interp, spl_thresh, step):
    """STFT -> per-bin projection -> ISTFT -> length alignment (train.py:38-66, :27-35), one fused kernel
    for min_max_freqs / max_phon and two passes for fletcher_munson."""
    x = L.f32c(p.detach().reshape(-1, p.shape[-1]))
    plan = L.plan_for(x, args)
    rows, T = x.shape
    frames = 1 + T // plan.hop
    out_len = plan.hop * (frames - 1) if clean_audio is None else int(clean_audio.shape[-1])
    out = torch.empty((rows, out_len), dtype=torch.float32, device=x.device)
    scratch = plan.scratch(rows, T)
    stream = L.stream_ptr(x.device)
    kind = args.norm_type
    if kind == "min_max_freqs":
        rc = L.lib.paa_project_min_max_freqs(plan.h, x.data_ptr(), out.data_ptr(), rows, T, out_len,
                                             float(args.min_freq_attack), float(args.max_freq_attack),
                                             L.step_ref(step), scratch, stream)
    elif kind == "max_phon":
        L.need_cuda(spl_thresh)
        thr = L.f32c(spl_thresh.detach().reshape(-1))
        if thr.numel() != plan.n_fft // 2 + 1:
            raise RuntimeError(f"spl_thresh has {thr.numel()} bins, expected {plan.n_fft // 2 + 1}")
        rc = L.lib.paa_project_max_phon(plan.h, x.data_ptr(), out.data_ptr(), rows, T, out_len, thr.data_ptr(),
                                        float(args.phon_reference_db), L.step_ref(step), scratch, stream)
    elif kind == "fletcher_munson":
        plan.set_fm_grid(interp)
        # pass B defaults to the identity form s*q (ISTFT(s*STFT(q)) = s*q on the reconstructed span, ~5e-7 from the
        # literal round trip); --fm_exact_roundtrip asks for the reference's second transform
        exact = 1 if getattr(args, "fm_exact_roundtrip", False) else 0
        rc = L.lib.paa_project_fletcher_munson(plan.h, x.data_ptr(), out.data_ptr(), rows, T, out_len,
                                               float(args.fm_epsilon), exact, L.step_ref(step), scratch, stream)
    else:
        raise ValueError(f"Unsupported frequency-domain norm_type: {kind!r}")
    L.check(rc, plan.h)
    return out.reshape(p.shape[:-1] + (out_len,))


def _dispatch(p, clean_audio, args, interp, spl_thresh, step):
    kind = args.norm_type
    if kind in FREQ_NORMS:
        return _frequency_domain(p, clean_audio, args, interp, spl_thresh, step)
    if kind == "l2":
        return projections.project_l2(p, args.l2_size, _step=step)
    if kind == "linf":
        return projections.project_linf(p, -args.linf_size, args.linf_size, _step=step)
    if kind == "snr":
        if clean_audio is None:
            raise ValueError("SNR projection requires clean_audio ro compare to")
        return projections.project_snr(clean=clean_audio, perturbation=p, snr_db=args.snr_db, _step=step)
    if kind == "tv":
        if clean_audio is None:
            raise ValueError("TV projection can benefit from clean_audio for bounds")
        return projections.project_tv(p=p, args=args, clean_audio=clean_audio, _step=step)
    raise ValueError(f"Unknown norm_type: {kind!r}")


def perturbation_constraint(p, clean_audio, args, interp, spl_thresh, parts=None):
    """Project perturbation p into the feasible set named by args.norm_type (train.py:69-99).
    Returns a new tensor; no autograd graph is attached.
    ``parts`` (mode U, training_utils/universal.py): clean statistics of the whole multi-rank batch for snr / tv."""
    L.need_cuda(p, clean_audio)
    with torch.no_grad():
        return _dispatch(p, clean_audio, args, interp, spl_thresh, L.make_step(L.STEP_NONE, parts=parts))


def step_and_project(p, grad, clean_audio, args, interp, spl_thresh, optimizer: Optional[torch.optim.Optimizer] = None,
                     parts=None):
    """One fused hot-path iteration: the optimiser step of train.py:155-175 on ``p`` given ``grad`` (= p.grad as
    autograd left it), then the projection -- in the same kernels, without host synchronisation.

    pgd : p + args.lr * sign(grad)                       (train.py:161)
    adam: torch.optim.Adam arithmetic on the optimiser's own state tensors (exp_avg, exp_avg_sq, step) and
          its current param_group lr, so StepLR and state_dict keep working (build.py:352-359).
    ``parts`` (mode U, ``UniversalExchange.publish``): per-rank partial gradients / clean statistics in peer memory;
          the kernels sum them while they step, ``grad`` is then only this rank's part.
    """
    L.need_cuda(p, grad, clean_audio)
    with torch.no_grad():
        g = L.f32c(grad.detach())
        if g.shape != p.shape:
            raise RuntimeError(f"grad shape {tuple(g.shape)} != p shape {tuple(p.shape)}")
        if args.optimizer_type == "pgd":
            step = L.make_step(L.STEP_PGD, g, args.lr, parts=parts)
        elif args.optimizer_type == "adam":
            if optimizer is None:
                raise ValueError("Adam optimizer selected but optimizer is None")
            step = optimizer.fused_step_descriptor(p, g)
            L.attach_parts(step, parts)
            out = _dispatch(p, clean_audio, args, interp, spl_thresh, step)
            optimizer.commit_step(step)               # only after the launch was accepted
            return out
        else:
            raise NotImplementedError(f"Optimization type not implemented: {args.optimizer_type!r}")
        return _dispatch(p, clean_audio, args, interp, spl_thresh, step)


def _chunked_forward_backward(model, clean_audio, p, target_texts, processor, args, micro, sign, loss_helpers):
    """train.py:136-145 + the backward of :158 / :170 over chunks of ``micro`` utterances.  Returns the whole-batch loss
    (sum of the chunk losses, detached) and the concatenated logits; p.grad holds the whole-batch gradient."""
    B = clean_audio.shape[0]
    per_row = p.shape[0] == B and B > 1
    total, logit_chunks = None, []
    for lo in range(0, B, micro):
        hi = min(lo + micro, B)
        pc = p[lo:hi] if per_row else p
        if getattr(args, "fused_compose", False):
            perturbed = compose.compose_clamp(clean_audio[lo:hi], pc)
        else:
            perturbed = (clean_audio[lo:hi] + pc).clamp_(-1.0, 1.0)
        loss_c, logits_c = loss_helpers.get_loss_for_training(model=model, data=perturbed, target_texts=target_texts[lo:hi],
                                                              processor=processor, args=args)
        (sign * loss_c).backward()
        total = loss_c.detach() if total is None else total + loss_c.detach()
        logit_chunks.append(logits_c.detach())
    return total, torch.cat(logit_chunks, dim=0)


def train_epoch(args, train_data_loader, p, model, epoch, processor, interp, wer_metric, spl_thresh, optimizer):
    """One epoch optimising a universal perturbation over the loader -- the loop of train.py:103-182 with the
    step + projection replaced by the fused call.  The wav2vec2 forward/backward stays on PyTorch."""
    try:
        from ..core import loss_helpers
    except ImportError:
        from core import loss_helpers
    ctc_scores, wer_scores, times = [], [], []
    defer = bool(getattr(args, "defer_metrics", False))     # not in the reference: no host sync inside the loop
    pending = []
    model.eval()
    logger.info("timestamp: %s | starting epoch: %d", datetime.now(), epoch)
    direction = +1 if args.attack_mode == "untargeted" else -1

    for clean_audio, target_texts in train_data_loader:
        t0 = time.perf_counter()
        clean_audio = clean_audio.to(args.device, non_blocking=True)
        clean_audio.requires_grad_(False)
        p.requires_grad_(True)
        if p.grad is not None:
            p.grad = None
        micro = int(getattr(args, "micro_batch", 0) or 0)
        chunked = 0 < micro < clean_audio.shape[0]
        if chunked:
            # the gradient source in chunks: forward + backward per chunk, p.grad accumulates the
            # chunk gradients (CTC reduction "sum": they add up to the whole-batch gradient); the step + projection
            # below sees the whole batch once
            if args.optimizer_type == "adam":
                if optimizer is None:
                    raise ValueError("Adam optimizer selected but optimizer is None")
                optimizer.zero_grad(set_to_none=True)
            sign = direction if args.optimizer_type == "pgd" else -1 * direction
            loss, logits = _chunked_forward_backward(model, clean_audio, p, target_texts, processor, args, micro, sign,
                                                     loss_helpers)
        else:
            if getattr(args, "fused_compose", False):          # one kernel forward, one backward (core/compose.py)
                perturbed = compose.compose_clamp(clean_audio, p)
            else:
                perturbed = (clean_audio + p).clamp_(-1.0, 1.0)
            loss, logits = loss_helpers.get_loss_for_training(model=model, data=perturbed, target_texts=target_texts,
                                                              processor=processor, args=args)
        if defer:
            # keep the loss and the greedy ids on the device; one synchronisation at the end of the epoch
            pending.append((loss.detach(), logits.detach().argmax(-1), target_texts))
        else:
            ctc_scores.append(float(loss.item()))
            with torch.inference_mode():
                wer = loss_helpers.compute_wer(logits=logits, target_texts=target_texts, processor=processor,
                                               wer_metric=wer_metric)
            wer_scores.append(float(wer))

        exchange = getattr(args, "universal_exchange", None)      # mode U: one perturbation shared by all ranks
        if args.optimizer_type == "pgd":
            if not chunked:
                (direction * loss).backward()
            parts = exchange.publish(p.grad, clean_audio, args.norm_type) if exchange is not None else None
            p = step_and_project(p, p.grad, clean_audio, args, interp, spl_thresh, parts=parts).detach()
        elif args.optimizer_type == "adam":
            if optimizer is None:
                raise ValueError("Adam optimizer selected but optimizer is None")
            if not chunked:
                optimizer.zero_grad(set_to_none=True)
                (-1 * direction * loss).backward()
            parts = exchange.publish(p.grad, clean_audio, args.norm_type) if exchange is not None else None
            with torch.no_grad():
                p.data = step_and_project(p.data, p.grad, clean_audio, args, interp, spl_thresh, optimizer=optimizer,
                                          parts=parts)
        else:
            raise NotImplementedError(f"Optimization type not implemented: {args.optimizer_type!r}")
        times.append(time.perf_counter() - t0)

    for loss_t, ids_t, texts in pending:                       # same numbers, computed after the last step was enqueued
        ctc_scores.append(float(loss_t.item()))
        wer_scores.append(float(loss_helpers.wer_from_ids(ids_t, texts, processor, wer_metric)))
    return TrainEpochResult(p=p, avg_ctc=_avg(ctc_scores), avg_wer=_avg(wer_scores))
