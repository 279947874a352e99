#!/usr/bin/env python
"""bench.py -- attack audio-sec/s per PGD step, and the projection's HBM GB/s against the measured peak.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one iteration of the reference's attack loop (src/training_utils/train.py:126-175) on one
synthetic batch: x_adv = clamp(clean + p) -> wav2vec2-base forward + CTC loss -> backward -> PGD step +
projection of p.  The workload is BASELINE.json configs[1]: targeted "delete" x5, snr 40 dB, batch 32 x 10 s of
16 kHz audio per GPU, random-init Wav2Vec2Config() weights, one perturbation row per utterance.  wav2vec2 and the
CTC loss are PyTorch/cuDNN (the untouched gradient source); the step + projection is libpaa.so.

`value`  : device-timed, inputs resident in HBM, transcripts/WER decoded after the timed region.
`e2e`    : the same metric through the public API with HOST buffers: each step copies the clean batch from pinned
           host memory, and reads back the loss and the greedy transcript ids (the reference loop's loss.item() and
           WER), inside the timed region.
`roofline`: the step+projection call timed with CUDA events inside the timed steps; achieved = algorithmic bytes
            / that time; peak = MEASURED_PEAKS.json hbm_gbs.
`cpu_baseline` / `--impl reference`: the CPU oracle port of the reference (oracle/paa_oracle.py, the one place a
           bench leg may execute it) on a bounded sample of the same workload.  The same leg also times the port as
           torch eager on this GPU per norm_type (`projection_sweep.*.torch_eager_ms`, SURVEY.md 2.2's
           kernel-for-kernel bar) and records the parity of the two results; `--no-cpu` skips all of it.
`--mode-u` (N > 1): ONE universal perturbation shared by all ranks instead of independent shards; the kernels sum the
           per-rank partial gradients from peer memory while they step (training_utils/universal.py).
N > 1: one process per GPU (torchrun), each rank attacks its own utterance shard with no collective on the hot
path; one NCCL all-reduce of the (edit errors, reference words) counters ends the run.  Weak scaling.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR = 16000
BATCH, SECONDS = 32, 10
SNR_DB = 40.0
LR = 1e-4
UNTARGETED_TEXT = "hello world this is a test"
METRIC = "attack audio-sec/s per PGD step"
# one `ncu --set full` capture of the dominant kernel at this workload (profiles/r01h_ncu_full.txt), bytes per launch
NCU_TRAFFIC_BYTES = 63_775_744


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def build_model(device):
    from transformers import Wav2Vec2Config, Wav2Vec2ForCTC
    torch.manual_seed(0)
    model = Wav2Vec2ForCTC(Wav2Vec2Config()).eval().to(device)
    return model


def synth(rank: int, batch: int, T: int, device, rows: int):
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    clean = (torch.rand(batch, T, generator=g) * 2 - 1) * 0.1
    p0 = torch.randn(rows, T, generator=g) * 0.01
    return clean, p0


def hot_bytes(rows: int, batch: int, T: int) -> int:
    """Algorithmic bytes of the PGD step + snr projection: pass A reads p, grad, writes q
    (12 B per element of p) and reads clean once (4 B per element of clean); pass B reads q, writes p (8 B)."""
    return 20 * rows * T + 4 * batch * T


# ---------------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch.distributed as dist
    import paa_b200
    from paa_b200 import paa_lib as L
    from paa_b200.core import loss_helpers
    from paa_b200.training_utils import parser as pparser

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    T = SECONDS * SR
    mode_u = bool(a.mode_u) and world > 1
    rows = 1 if (a.universal or mode_u) else BATCH
    args = pparser.create_arg_parser().parse_args(
        ["--norm_type", "snr", "--snr_db", str(SNR_DB), "--optimizer_type", "pgd", "--attack_mode", "targeted",
         "--lr", str(LR)])
    args.device = str(dev)
    model = build_model(dev)
    for q in model.parameters():          # weight gradients are never used by the attack
        q.requires_grad_(False)
    clean_h, p0 = synth(rank, BATCH, T, dev, rows)
    clean_h = clean_h.pin_memory()
    clean_d = clean_h.to(dev)
    texts = [UNTARGETED_TEXT] * BATCH
    labels = loss_helpers.encode_labels(
        loss_helpers.clean_transcripts([" ".join([args.target] * args.target_reps)] * BATCH), dev)
    exch = None
    if mode_u:
        # SURVEY.md 8e mode U: ONE (1,T) perturbation shared by all ranks; the kernels sum the per-rank partial
        # gradients and clean statistics from peer memory while they step (training_utils/universal.py)
        from paa_b200.training_utils import universal
        exch = universal.UniversalExchange(1, T, dev, backend=a.mode_u_backend)
        p0 = universal.broadcast_perturbation(p0.to(dev))
        p = paa_b200.perturbation_constraint(p0, clean_d, args, None, None, parts=exch.publish(None, clean_d))
        args.universal_exchange = exch
    else:
        p = paa_b200.perturbation_constraint(p0.to(dev), clean_d, args, None, None)
    direction = -1.0                      # targeted: descend the loss (train.py:124)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]

    def one_step(p, clean, k=None):
        p = p.detach().requires_grad_(True)
        x_adv = (clean + p).clamp_(-1.0, 1.0)
        out = model(input_values=x_adv, labels=labels)
        (direction * out.loss).backward()
        if k is not None:
            ev[k][0].record()
        parts = exch.publish(p.grad, clean, args.norm_type) if exch is not None else None
        p_new = paa_b200.step_and_project(p.detach(), p.grad, clean, args, None, None, parts=parts)
        if k is not None:
            ev[k][1].record()
        return p_new, out.loss.detach(), out.logits.detach().argmax(-1)

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- value: inputs resident, bookkeeping after the timed region ---------------------------------------
    for _ in range(a.warmup):
        p, _, _ = one_step(p, clean_d)
    fence()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = L.lib.paa_launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    losses, ids = [], []
    t0.record()
    for k in range(a.steps):
        p, loss, pred = one_step(p, clean_d, k)
        losses.append(loss)
        ids.append(pred)
    wer = loss_helpers.WerMetric()
    counters = torch.zeros(2, dtype=torch.int64, device=dev)
    t1.record()
    fence()
    launches = L.lib.paa_launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
    proj_ms = statistics.mean(e0.elapsed_time(e1) for e0, e1 in ev)
    # transcripts / WER of the last step, counters summed across ranks by the one collective of the run
    hyp = [t.lower() for t in loss_helpers.greedy_decode(ids[-1])]
    ref = loss_helpers.clean_transcripts([" ".join([args.target] * args.target_reps)] * BATCH)
    wer.compute(predictions=hyp, references=ref)
    counters += torch.tensor([wer.errors, wer.words], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    ms_per_step = float(ms) / a.steps
    value = world * BATCH * SECONDS / (ms_per_step / 1e3)

    # ---- e2e: the public API a user of the reference calls -- train_epoch over a loader of HOST batches ---------
    # Every step: H2D of the pinned clean batch, compose, wav2vec2 + CTC, loss.item(), argmax ids D2H + greedy
    # decode + WER counters (libpaa), backward, fused step + projection (train.py:126-175).
    from paa_b200.training_utils import train as ptrain
    ids_bytes = ids[-1].numel() * ids[-1].element_size()

    def e2e_epoch(pp, n_steps):
        loader = [(clean_h, texts)] * n_steps
        res = ptrain.train_epoch(args, loader, pp.detach(), model, 0, None, None, loss_helpers.WerMetric(), None, None)
        return res.p.detach()

    e2e_p = e2e_epoch(p.detach().clone(), max(1, a.warmup // 2))
    fence()
    w0 = time.perf_counter()
    e2e_p = e2e_epoch(e2e_p, a.steps)
    fence()
    e2e_ms = torch.tensor([(time.perf_counter() - w0) * 1e3], device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * BATCH * SECONDS / (float(e2e_ms) / a.steps / 1e3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak()
    nbytes = hot_bytes(rows, BATCH, T)
    achieved = nbytes / (proj_ms * 1e-3) / 1e9
    sweep = projection_sweep(dev) if (world == 1 and not a.no_sweep) else None
    cpu = cpu_baseline(a, sample_batch=2) if (world == 1 and not a.no_cpu) else None
    if cpu and sweep:
        baseline_torch_eager(dev, sweep)
    line = {
        "metric": METRIC, "value": round(value, 2), "unit": "audio-s/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: targeted 'delete'x5, snr 40 dB, PGD, batch 32 x 10 s @16 kHz per GPU, "
                               "random-init wav2vec2-base" + (", ONE universal (1,T) perturbation shared by all ranks (mode U)"
                                                             if mode_u else ", universal (1,T) perturbation" if a.universal
                                                             else ", one perturbation row per utterance"),
                   "batch_per_gpu": BATCH, "seconds": SECONDS, "p_rows": rows,
                   "parallelism": (f"universal-dp x{world} ({a.mode_u_backend}: partial gradients summed from peer memory "
                                   "inside the step kernel)" if mode_u else f"utterance-sharded x{world}"),
                   "l2_between_iters": "working set per step (activations, GBs) exceeds the 126 MB L2"},
        "e2e": {"value": round(e2e_value, 2), "unit": "audio-s/s", "h2d_bytes_per_step": clean_h.numel() * 4,
                "d2h_bytes_per_step": ids_bytes + 4,
                "api": "paa_b200.training_utils.train.train_epoch (mirror of train.py:103-182), wall clock"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_fused<snr,pgd>: PGD step + energy reduce + grid barrier + rescale, one cooperative launch",
                     "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                     "traffic": NCU_TRAFFIC_BYTES if rows == BATCH else None, "algorithmic_bytes": nbytes,
                     "traffic_source": "profiles/r01h_ncu_full.txt: dram__bytes_read.sum + dram__bytes_write.sum of "
                                       "k_fused<1,1> (61.49 + 2.29 MB; the 20.5 MB result is still in L2 at kernel end)", "avg_call_us": round(proj_ms * 1e3, 2),
                     "peak_source": peak_src},
        "wer_counters": {"errors": int(counters[0]), "ref_words": int(counters[1])},
        "loss_last": round(float(losses[-1]), 3),
    }
    if cpu:
        line["cpu_baseline"] = cpu
    if sweep:
        line["projection_sweep"] = sweep
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


SWEEP_CASES = [("linf", 4, 5, 1e-3, 12), ("snr", 32, 10, 0.01, 24), ("fletcher_munson", 64, 15, 0.1, 20),
               ("fletcher_munson+identity", 64, 15, 0.1, 20), ("max_phon", 64, 15, 0.03, 12), ("tv", 128, 10, 0.01, 24),
               ("min_max_freqs", 128, 10, 0.01, 12), ("l2", 512, 10, 0.01, 20),
               # Adam instead of PGD: the step alone is 28 B/elem (R p,g,m,v; W p,m,v) in place of 12
               ("linf+adam", 128, 10, 1e-3, 28), ("snr+adam", 32, 10, 0.01, 40), ("max_phon+adam", 64, 15, 0.03, 36)]


def sweep_inputs(dev, B, sec, sigma):
    T = sec * SR
    g = torch.Generator(device=dev).manual_seed(1234)
    clean = (torch.rand(B, T, generator=g, device=dev) * 2 - 1) * 0.1
    p = torch.randn(B, T, generator=g, device=dev) * sigma
    grad = torch.randn(B, T, generator=g, device=dev)
    return clean, p, grad


def projection_sweep(dev, iters: int = 20):
    """Step + projection alone at BASELINE.json's shapes, per-utterance rows, CUDA events, L2 flushed by the
    working set being larger than L2 where it is (noted per entry)."""
    import paa_b200
    from paa_b200.core import iso
    from paa_b200.training_utils import build as pbuild, parser as pparser
    peak, _ = measured_peak()
    interp = iso.build_weight_interpolator()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    out = {}
    for name, B, sec, sigma, bpe in SWEEP_CASES:
        norm = name.split("+")[0]
        T = sec * SR
        clean, p, grad = sweep_inputs(dev, B, sec, sigma)
        args = pparser.create_arg_parser().parse_args(["--norm_type", norm, "--optimizer_type", "pgd", "--snr_db", "40"])
        args.device = str(dev)
        # "+identity": pass B of fletcher_munson as s*q (ISTFT(s*STFT(q)) = s*q) instead of the literal round trip
        args.fm_identity_roundtrip = name.endswith("+identity")
        thr = pbuild.init_phon_threshold_tensor(args)
        opt = None
        if name.endswith("+adam"):
            args.optimizer_type = "adam"
            opt, _ = pbuild.create_optimizer(args, p)
        for _ in range(3):
            paa_b200.step_and_project(p, grad, clean, args, interp, thr, optimizer=opt)
        times = []
        for _ in range(iters):
            flush.zero_()
            torch.cuda._sleep(400_000)            # keep the GPU busy while the host enqueues: no launch latency in the interval
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            paa_b200.step_and_project(p, grad, clean, args, interp, thr, optimizer=opt)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = statistics.median(times)
        gbs = bpe * B * T / (ms * 1e-3) / 1e9
        out[name] = {"shape": f"{B}x{sec}s", "bytes_per_elem": bpe, "ms": round(ms, 4), "GB/s": round(gbs, 1),
                     "frac_of_measured_peak": round(gbs / peak, 4), "audio_s_per_s": round(B * sec / (ms * 1e-3), 1)}
        del clean, p, grad
        torch.cuda.empty_cache()
    out.update(compose_sweep(dev, flush, peak, iters))
    return out


def compose_sweep(dev, flush, peak, iters):
    """The input side: x_adv = clamp(clean + p) and dL/dp, universal (1,T) and per-utterance p,
    batch 32 x 10 s; torch eager (add, clamp_, autograd's mask + batch sum) timed beside it."""
    from paa_b200.core.compose import compose_clamp
    B, T = BATCH, SECONDS * SR
    g = torch.Generator(device=dev).manual_seed(7)
    clean = (torch.rand(B, T, generator=g, device=dev) * 2 - 1) * 0.9
    w = torch.randn(B, T, generator=g, device=dev)
    res = {}
    for rows in (1, B):
        p = (torch.randn(rows, T, generator=g, device=dev) * 0.3).requires_grad_(True)

        def timed(fn):
            ts = []
            for _ in range(iters):
                flush.zero_()
                torch.cuda._sleep(400_000)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            return statistics.median(ts), r

        fwd_ms, x = timed(lambda: compose_clamp(clean, p))
        bwd_ms, _ = timed(lambda: torch.autograd.grad(x, p, w, retain_graph=True))
        tf_ms, xt = timed(lambda: (clean + p).clamp_(-1.0, 1.0))
        tb_ms, _ = timed(lambda: torch.autograd.grad(xt, p, w, retain_graph=True))
        fb, bb = 8 * B * T + 4 * rows * T, 8 * B * T + 8 * rows * T       # algorithmic bytes: fwd R clean,p W x; bwd R clean,g,p W gp
        tag = "universal" if rows == 1 else "per_utterance"
        res[f"compose_fwd_{tag}"] = {"shape": f"{B}x{SECONDS}s", "ms": round(fwd_ms, 4), "GB/s": round(fb / fwd_ms / 1e6, 1),
                                     "frac_of_measured_peak": round(fb / fwd_ms / 1e6 / peak, 4), "torch_eager_ms": round(tf_ms, 4)}
        res[f"compose_bwd_{tag}"] = {"shape": f"{B}x{SECONDS}s", "ms": round(bwd_ms, 4), "GB/s": round(bb / bwd_ms / 1e6, 1),
                                     "frac_of_measured_peak": round(bb / bwd_ms / 1e6 / peak, 4), "torch_eager_ms": round(tb_ms, 4)}
    return res


# ---------------------------------------------------------------------------------------------------------
def oracle_step_seconds(batch: int, steps: int, warmup: int, rows: int):
    """The CPU oracle port of the reference's attack iteration on `batch` x 10 s; returns seconds per step."""
    from oracle import paa_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    from transformers import Wav2Vec2Config, Wav2Vec2ForCTC
    torch.manual_seed(0)
    model = Wav2Vec2ForCTC(Wav2Vec2Config()).eval()
    T = SECONDS * SR
    clean, p = synth(0, batch, T, "cpu", rows)
    hp = orc.Hyper(norm_type="snr", snr_db=SNR_DB, optimizer_type="pgd", attack_mode="targeted", lr=LR)
    p = orc.constrain(p, clean, hp)
    for _ in range(warmup):
        p, _, _ = orc.attack_iteration(model, p, clean, [UNTARGETED_TEXT] * batch, hp)
    t0 = time.perf_counter()
    for _ in range(steps):
        p, loss, _ = orc.attack_iteration(model, p, clean, [UNTARGETED_TEXT] * batch, hp)
    return (time.perf_counter() - t0) / steps


def baseline_torch_eager(dev, sweep):
    """Part of the baseline leg (the only place the oracle port runs): the same torch ops the reference executes,
    eager, on THIS GPU -- the kernel-for-kernel bar of SURVEY.md 2.2 -- per norm at the sweep's shapes and seeds, written
    next to our numbers together with the parity of the two results.  fletcher_munson includes the reference's
    D2H -> host bilinear interpolation -> H2D round trip."""
    import paa_b200
    from oracle import paa_oracle as orc
    from paa_b200.core import iso
    from paa_b200.training_utils import build as pbuild, parser as pparser
    interp = iso.build_weight_interpolator()
    it_cpu = orc.build_weight_interpolator()
    for name, B, sec, sigma, _ in SWEEP_CASES:
        if name not in sweep or name.endswith("+adam"):
            continue
        norm = name.split("+")[0]
        try:
            clean, p, grad = sweep_inputs(dev, B, sec, sigma)
            args = pparser.create_arg_parser().parse_args(["--norm_type", norm, "--optimizer_type", "pgd", "--snr_db", "40"])
            args.device = str(dev)
            args.fm_identity_roundtrip = name.endswith("+identity")
            thr = pbuild.init_phon_threshold_tensor(args)
            hp = orc.Hyper(norm_type=norm, optimizer_type="pgd", snr_db=SNR_DB)
            reps = 2 if norm == "fletcher_munson" else 5
            want = orc.step_and_constrain(p, grad, clean, hp, it_cpu, thr)
            torch.cuda.synchronize()
            w0 = time.perf_counter()
            for _ in range(reps):
                orc.step_and_constrain(p, grad, clean, hp, it_cpu, thr)
            torch.cuda.synchronize()
            eager_ms = (time.perf_counter() - w0) * 1e3 / reps
            got = paa_b200.step_and_project(p, grad, clean, args, interp, thr)
            sweep[name]["torch_eager_ms"] = round(eager_ms, 3)
            sweep[name]["speedup_vs_torch_eager"] = round(eager_ms / sweep[name]["ms"], 1)
            sweep[name]["max_rel_err_vs_torch_eager"] = float(f"{float((got - want).abs().max() / want.abs().max()):.2e}")
            del want, got, clean, p, grad
        except Exception as exc:                                  # the comparator must never break the bench line
            sweep[name]["torch_eager_ms"] = f"unavailable: {type(exc).__name__}"
        torch.cuda.empty_cache()


def cpu_baseline(a, sample_batch: int):
    sec = oracle_step_seconds(sample_batch, steps=1, warmup=1, rows=1 if a.universal else sample_batch)
    return {"value": round(sample_batch * SECONDS / sec, 3), "unit": "audio-s/s", "cores": torch.get_num_threads(),
            "kind": "port", "sample": f"oracle/paa_oracle.py attack_iteration, batch {sample_batch} x {SECONDS} s "
                                      f"(1/{BATCH // sample_batch} of the step's batch), 1 warm-up + 1 timed step"}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 2
    rows = 1 if a.universal else sample
    sec = oracle_step_seconds(sample, steps=a.steps, warmup=a.warmup, rows=rows)
    value = sample * SECONDS / sec
    world = int(os.environ.get("WORLD_SIZE", "1"))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": "audio-s/s", "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(sec * 1e3, 2), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: targeted 'delete'x5, snr 40 dB, PGD, 10 s @16 kHz, random-init wav2vec2-base; "
                               f"each step is a bounded sample of batch {sample} of the 32-utterance batch",
                   "batch_per_gpu": BATCH, "seconds": SECONDS, "p_rows": rows},
        "cpu_baseline": {"value": round(value, 3), "unit": "audio-s/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"oracle/paa_oracle.py attack_iteration, batch {sample} x {SECONDS} s per step"},
        "e2e": {"value": round(value, 3), "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--universal", action="store_true", help="one (1,T) perturbation shared by the batch, as the reference's loop")
    ap.add_argument("--mode-u", dest="mode_u", action="store_true",
                    help="N>1: one universal perturbation shared by all ranks (SURVEY.md 8e mode U) instead of independent shards")
    ap.add_argument("--mode-u-backend", dest="mode_u_backend", choices=["symmetric", "nccl"], default="symmetric")
    ap.add_argument("--no-cpu", dest="no_cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sweep", dest="no_sweep", action="store_true", help="skip the per-norm projection sweep")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
