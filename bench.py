#!/usr/bin/env python
"""bench.py -- attack audio-sec/s per PGD step, and the projection's HBM GB/s against the measured peak.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one iteration of the reference's attack loop (src/training_utils/train.py:126-175) on one synthetic
batch: x_adv = clamp(clean + p) -> wav2vec2-base forward + CTC loss -> backward -> PGD step + projection of p.

Workload = BASELINE.json configs[4], the configuration the 1/2/4/8-GPU metric is quoted on: untargeted l2 attack,
batch 512 x 10 s of 16 kHz audio IN TOTAL, utterance-sharded (512/N rows per rank, strong scaling), random-init
Wav2Vec2Config() weights, one perturbation row per utterance, NCCL all-reduce of the WER counters.  wav2vec2 and the CTC
loss are PyTorch/cuDNN (the untouched gradient source), called in chunks of --micro-batch utterances (CTC reduction
"sum": chunk gradients add up); the step + projection is ONE libpaa launch over the whole shard.

`value`  : device-timed, inputs resident in HBM, transcripts/WER decoded after the timed region.
`e2e`    : the same metric through the public API (training_utils.train.train_epoch) with HOST buffers: each step
           copies the clean shard from pinned host memory and reads back the loss and the greedy transcript ids (the
           reference loop's loss.item() and WER), inside the timed region.
`roofline`: the step+projection call (k_fused<l2,pgd>) timed with CUDA events inside the timed steps; achieved =
           algorithmic bytes (20 B per element of p) / that time; peak = MEASURED_PEAKS.json.
N = 1 only, after the timed legs:
`configs1`, `universal`: the same loop on configs[1] (targeted snr 40 dB, 32 x 10 s) and on configs[4] with the
           reference's universal (1,T) perturbation, a few steps each.
`projection_sweep`: step + projection alone for every norm_type at BASELINE.json's shapes (HBM GB/s, and fp32 GFLOP/s
           for the STFT family), device time and host wall clock per call.
`cpu_baseline` / `--impl reference`: the CPU oracle port of the reference (oracle/paa_oracle.py, the one place a
           bench leg may execute it) on a bounded sample of the same workload.  The same leg times the port as torch
           eager on this GPU per norm_type (`projection_sweep.*.torch_eager_ms`, SURVEY.md 2.2's kernel-for-kernel
           bar) with the parity of the two results, and runs `transcript_identity` (oracle/transcript_check.py):
           reference arithmetic vs libpaa on wav2vec2-base, 20 steps; `--no-cpu` skips all of it.
N > 1: one process per GPU (torchrun), each rank attacks its own utterance shard with no collective on the hot path; one
NCCL all-reduce of the (edit errors, reference words) counters ends the run.  `mode_u` (N > 1): ONE universal
perturbation shared by all ranks -- publish + step + projection per step, peer-memory kernels vs NCCL all-reduce, and
the bit-identity of the result across ranks.
"""
from __future__ import annotations

import argparse
import json
import math
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
TOTAL_BATCH, SECONDS = 512, 10           # BASELINE.json configs[4]
NORM = "l2"
C1_BATCH, C1_SNR_DB = 32, 40.0           # configs[1], secondary leg
LR = 1e-4
UNTARGETED_TEXT = "hello world this is a test"
METRIC = "attack audio-sec/s per PGD step"
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # 74.4: CUDA-core fp32 FMA peak at the boost clock
WORKLOAD = ("configs[4]: untargeted l2 attack (PGD), batch 512 x 10 s @16 kHz in total, utterance-sharded, "
            "random-init wav2vec2-base")
# `ncu --set full` of the dominant kernel at the N=1 workload (profiles/), bytes per launch; None until captured
NCU_TRAFFIC_BYTES = 895_134_208 + 540_193_792
NCU_TRAFFIC_SOURCE = ("profiles/r02i_ncu_full_k_fused_final.txt: dram__bytes_read.sum + dram__bytes_write.sum of k_fused<l2,pgd> at "
                      "512 x 10 s (895.1 + 540.2 MB; below the algorithmic 1638.4 MB because 12 M of the 82 M stepped elements "
                      "stay in registers / shared memory across the grid barrier and another ~12 M are re-read from L2)")


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


def hot_bytes(norm: str, rows: int, batch: int, T: int) -> int:
    """Algorithmic bytes of the PGD step + projection.  l2: pass A reads p, grad, writes q
    (12 B per element of p), pass B reads q, writes p (8 B).  snr adds one read of clean (4 B per element of clean)."""
    return 20 * rows * T + (4 * batch * T if norm in ("snr", "tv") else 0)


class AttackLoop:
    """The reference's loop (train.py:126-175) on device-resident inputs: compose -> wav2vec2 + CTC in chunks of
    `micro` utterances -> backward (p.grad accumulates the chunk gradients) -> ONE fused step + projection over all rows."""

    def __init__(self, model, clean, labels, args, micro, direction, exch=None):
        self.model, self.clean, self.labels, self.args = model, clean, labels, args
        self.micro = micro if micro and micro < clean.shape[0] else clean.shape[0]
        self.direction, self.exch = direction, exch
        self.events = []

    def step(self, p, timed=False):
        import paa_b200
        B = self.clean.shape[0]
        per_row = p.shape[0] == B and B > 1
        p = p.detach().requires_grad_(True)
        loss, ids = None, []
        for lo in range(0, B, self.micro):
            hi = min(lo + self.micro, B)
            x_adv = (self.clean[lo:hi] + (p[lo:hi] if per_row else p)).clamp_(-1.0, 1.0)
            out = self.model(input_values=x_adv, labels=self.labels[lo:hi])
            (self.direction * out.loss).backward()
            loss = out.loss.detach() if loss is None else loss + out.loss.detach()
            ids.append(out.logits.detach().argmax(-1))
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        parts = self.exch.publish(p.grad, self.clean, self.args.norm_type) if self.exch is not None else None
        p_new = paa_b200.step_and_project(p.detach(), p.grad, self.clean, self.args, None, None, parts=parts)
        if timed:
            e1.record()
            self.events.append((e0, e1))
        return p_new, loss, torch.cat(ids, 0)

    def proj_ms(self):
        return statistics.mean(e0.elapsed_time(e1) for e0, e1 in self.events)


def make_args(norm, mode, dev, micro=0, **over):
    from paa_b200.training_utils import parser as pparser
    args = pparser.create_arg_parser().parse_args(["--norm_type", norm, "--optimizer_type", "pgd", "--attack_mode", mode,
                                                   "--lr", str(LR), "--micro_batch", str(micro)])
    for k, v in over.items():
        setattr(args, k, v)
    args.device = str(dev)
    return args


def labels_for(args, batch, dev):
    from paa_b200.core import loss_helpers
    texts = [UNTARGETED_TEXT] * batch
    if args.attack_mode == "targeted":
        texts = [" ".join([args.target] * args.target_reps)] * batch
    return loss_helpers.encode_labels(loss_helpers.clean_transcripts(texts), dev), loss_helpers.clean_transcripts(texts)


# ---------------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch.distributed as dist
    import paa_b200
    from paa_b200 import paa_lib as L
    from paa_b200.core import loss_helpers
    from paa_b200.training_utils import sharding, train as ptrain

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
    lo, hi = sharding.shard_bounds(TOTAL_BATCH, rank, world)       # this rank's utterances
    batch = hi - lo
    rows = 1 if a.universal else batch
    args = make_args(NORM, "untargeted", dev, micro=a.micro_batch)
    model = build_model(dev)
    for q in model.parameters():          # weight gradients are never used by the attack
        q.requires_grad_(False)
    clean_h, p0 = synth(rank, batch, T, dev, rows)
    clean_h = clean_h.pin_memory()
    clean_d = clean_h.to(dev)
    labels, ref_texts = labels_for(args, batch, dev)
    p = paa_b200.perturbation_constraint(p0.to(dev), clean_d, args, None, None)
    loop = AttackLoop(model, clean_d, labels, args, a.micro_batch, +1.0)

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- value: inputs resident, bookkeeping after the timed region ---------------------------------------
    for _ in range(a.warmup):
        p, _, _ = loop.step(p)
    fence()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = L.lib.paa_launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    losses, ids = [], []
    t0.record()
    for k in range(a.steps):
        p, loss, pred = loop.step(p, timed=True)
        losses.append(loss)
        ids.append(pred)
    t1.record()
    fence()
    launches = L.lib.paa_launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
    proj_ms = loop.proj_ms()
    # transcripts / WER of the last step, counters summed across ranks by the one collective of the run
    hyp = [t.lower() for t in loss_helpers.greedy_decode(ids[-1])]
    e_cnt, w_cnt, _ = sharding.allreduce_wer(ref_texts, hyp, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms) / a.steps
    value = TOTAL_BATCH * SECONDS / (ms_per_step / 1e3)

    # ---- e2e: the public API a user of the reference calls -- train_epoch over a loader of HOST batches ---------
    # Every step: H2D of the pinned clean shard, compose, wav2vec2 + CTC (chunked), loss.item(), argmax ids D2H + greedy
    # decode + WER counters (libpaa), backward, fused step + projection (train.py:126-175).
    texts = [UNTARGETED_TEXT] * batch
    ids_bytes = ids[-1].numel() * ids[-1].element_size()

    def e2e_epoch(pp, n_steps, run_args=None):
        loader = [(clean_h, texts)] * n_steps
        res = ptrain.train_epoch(run_args or args, loader, pp.detach(), model, 0, None, None, loss_helpers.WerMetric(), None, None)
        return res.p.detach()

    e2e_p = e2e_epoch(p.detach().clone(), max(1, a.warmup // 2))
    fence()
    w0 = time.perf_counter()
    e2e_p = e2e_epoch(e2e_p, a.steps)
    fence()
    e2e_ms = torch.tensor([(time.perf_counter() - w0) * 1e3], device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = TOTAL_BATCH * SECONDS / (float(e2e_ms) / a.steps / 1e3)
    # the same epoch with the two opt-in pieces of the path's input side: the fused compose + clamp
    # kernels (core/compose.py) and the loss / transcript read-back deferred to the end of the epoch (no host sync per step)
    import copy
    fargs = copy.copy(args)
    fargs.fused_compose, fargs.defer_metrics = True, True
    f_steps = max(3, a.steps // 4)
    e2e_epoch(e2e_p.clone(), 1, fargs)
    fence()
    w0 = time.perf_counter()
    e2e_epoch(e2e_p.clone(), f_steps, fargs)
    fence()
    e2e_f_ms = torch.tensor([(time.perf_counter() - w0) * 1e3], device=dev)
    if world > 1:
        dist.all_reduce(e2e_f_ms, op=dist.ReduceOp.MAX)
    e2e_fused_value = TOTAL_BATCH * SECONDS / (float(e2e_f_ms) / f_steps / 1e3)

    mode_u = mode_u_record(dev, rank, world, clean_d, T) if (world > 1 and not a.no_mode_u) else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak()
    nbytes = hot_bytes(NORM, rows, batch, T)
    achieved = nbytes / (proj_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": round(value, 2), "unit": "audio-s/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD + (", universal (1,T) perturbation" if a.universal else ", one perturbation row per utterance"),
                   "total_batch": TOTAL_BATCH, "batch_per_gpu": batch, "seconds": SECONDS, "p_rows": rows,
                   "micro_batch": loop.micro, "norm_type": NORM, "optimizer": "pgd",
                   "parallelism": f"utterance-sharded x{world} (mode R: no collective on the hot path, one NCCL all-reduce of the WER counters)",
                   "l2_between_iters": "working set per step (activations, GBs) exceeds the 126 MB L2"},
        "e2e": {"value": round(e2e_value, 2), "unit": "audio-s/s", "h2d_bytes_per_step": clean_h.numel() * 4,
                "d2h_bytes_per_step": ids_bytes + 4,
                "api": "paa_b200.training_utils.train.train_epoch (mirror of train.py:103-182, --micro_batch), wall clock",
                "fused_compose_defer_metrics": {"value": round(e2e_fused_value, 2), "steps": f_steps,
                                                "what": "the same epoch with --fused_compose --defer_metrics (libpaa compose + clamp forward / backward, "
                                                        "loss and transcripts read back once at the end of the epoch)"}},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_fused<l2,pgd>: PGD step + sum of squares + grid barrier + rescale, one cooperative launch",
                     "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                     "traffic": NCU_TRAFFIC_BYTES if (world == 1 and not a.universal) else None,
                     "algorithmic_bytes": nbytes, "bytes_per_elem": 20, "elements": rows * T,
                     "traffic_source": NCU_TRAFFIC_SOURCE, "avg_call_us": round(proj_ms * 1e3, 2),
                     "frac_of_spec_8TBs": round(achieved / 8000.0, 4), "peak_source": peak_src},
        "wer_counters": {"errors": int(e_cnt), "ref_words": int(w_cnt), "allreduce": "NCCL int64[2] sum" if world > 1 else "single rank"},
        "loss_last": round(float(losses[-1]), 3),
    }
    if mode_u:
        line["mode_u"] = mode_u
    if world == 1:
        # release the big buffers before the secondary legs
        del loop, clean_d, e2e_p
        torch.cuda.empty_cache()
        if not a.no_secondary:
            line["configs1"] = secondary_leg(model, dev, "snr", "targeted", C1_BATCH, C1_BATCH, steps=5, micro=0, snr_db=C1_SNR_DB)
            line["universal"] = secondary_leg(model, dev, NORM, "untargeted", TOTAL_BATCH, 1, steps=3, micro=a.micro_batch)
        sweep = projection_sweep(dev) if not a.no_sweep else None
        cpu = cpu_baseline(a) if not a.no_cpu else None
        if cpu and sweep:
            baseline_torch_eager(dev, sweep)
        if cpu:
            line["cpu_baseline"] = cpu
            line["transcript_identity"] = transcript_identity(model, dev)
        if sweep:
            line["projection_sweep"] = sweep
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def secondary_leg(model, dev, norm, mode, batch, rows, steps, micro, **over):
    """The attack loop on another BASELINE configuration / perturbation shape, a few device-timed steps (N = 1)."""
    import paa_b200
    T = SECONDS * SR
    args = make_args(norm, mode, dev, micro=micro, **over)
    clean, p0 = synth(0, batch, T, dev, rows)
    clean = clean.to(dev)
    labels, _ = labels_for(args, batch, dev)
    p = paa_b200.perturbation_constraint(p0.to(dev), clean, args, None, None)
    loop = AttackLoop(model, clean, labels, args, micro, +1.0 if mode == "untargeted" else -1.0)
    for _ in range(3):
        p, _, _ = loop.step(p)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        p, loss, _ = loop.step(p, timed=True)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / steps
    peak, _ = measured_peak()
    nbytes = hot_bytes(norm, rows, batch, T)
    gbs = nbytes / (loop.proj_ms() * 1e-3) / 1e9
    return {"workload": f"{mode} {norm}, PGD, batch {batch} x {SECONDS} s, p_rows {rows}", "steps": steps,
            "value": round(batch * SECONDS / (ms / 1e3), 2), "unit": "audio-s/s", "ms_per_step": round(ms, 3),
            "step_projection_us": round(loop.proj_ms() * 1e3, 2), "algorithmic_bytes": nbytes,
            "GB/s": round(gbs, 1), "frac_of_measured_peak": round(gbs / peak, 4), "loss_last": round(float(loss), 3)}


def mode_u_record(dev, rank, world, clean, T, iters: int = 30):
    """SURVEY.md 8e mode U on the ranks of this run: ONE universal (1,T) perturbation, every rank holds its own
    utterance shard.  Per step: publish (copy of the partial gradient, clean statistics, device-side barrier) + step +
    projection with the kernels summing the G partials from peer memory, against NCCL all-reduce + the ordinary
    kernels.  Microseconds per step (CUDA events, max over ranks) and the bit-identity of p across ranks."""
    import torch.distributed as dist
    import paa_b200
    from paa_b200.training_utils import universal
    out = {"what": f"publish + step + projection per step, universal (1,{T}) p, {clean.shape[0]} x {T} clean per rank, {world} GPUs",
           "us_per_step": {}, "bit_identical_across_ranks": True}
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    grad = torch.randn(1, T, generator=g, device=dev)
    for norm in ("l2", "snr"):
        args = make_args(norm, "untargeted", dev, snr_db=40.0)
        for backend in ("symmetric", "nccl"):
            try:
                p = universal.broadcast_perturbation(torch.randn(1, T, device=dev) * 0.01)
                exch = universal.UniversalExchange(1, T, dev, backend=backend)
                times = []
                for it in range(iters + 5):
                    dist.barrier()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    parts = exch.publish(grad, clean, norm)
                    p = paa_b200.step_and_project(p, grad, clean, args, None, None, parts=parts)
                    e1.record()
                    torch.cuda.synchronize()
                    if it >= 5:
                        times.append(e0.elapsed_time(e1))
                t = torch.tensor([statistics.median(times)], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                # every rank must hold the same bits: all-gather an integer checksum of the fp32 words
                chk = p.view(torch.int32).to(torch.int64).sum().reshape(1)
                allc = [torch.zeros_like(chk) for _ in range(world)]
                dist.all_gather(allc, chk)
                same = all(int(c) == int(allc[0]) for c in allc)
                out["us_per_step"][f"{norm}/{backend}"] = round(float(t) * 1e3, 1)
                out["bit_identical_across_ranks"] = out["bit_identical_across_ranks"] and same
            except Exception as exc:                                   # the record must never break the bench line
                out["us_per_step"][f"{norm}/{backend}"] = f"unavailable: {type(exc).__name__}: {exc}"[:160]
    return out


# ---------------------------------------------------------------------------------------------------------
# (name, norm, B, seconds, rows (0 = B), sigma, optimiser, algorithmic bytes per element of p [+ clean bytes per clean element])
SWEEP_CASES = [
    ("linf", 4, 5, 0, 1e-3, "pgd"), ("linf_128x10s", 128, 10, 0, 1e-3, "pgd"), ("snr", 32, 10, 0, 0.01, "pgd"),
    ("fletcher_munson", 64, 15, 0, 0.1, "pgd"), ("fletcher_munson+exact", 64, 15, 0, 0.1, "pgd"),
    ("max_phon", 64, 15, 0, 0.03, "pgd"), ("tv", 128, 10, 0, 0.01, "pgd"), ("min_max_freqs", 128, 10, 0, 0.01, "pgd"),
    ("l2", 512, 10, 0, 0.01, "pgd"),
    # the reference's own perturbation shape: ONE universal (1,T) row against the whole batch
    ("l2_universal", 512, 10, 1, 0.01, "pgd"), ("snr_universal", 32, 10, 1, 0.01, "pgd"), ("max_phon_universal", 64, 15, 1, 0.03, "pgd"),
    # Adam instead of PGD: the step alone is 28 B/elem (R p,g,m,v; W p,m,v) in place of 12
    ("linf+adam", 128, 10, 0, 1e-3, "adam"), ("snr+adam", 32, 10, 0, 0.01, "adam"), ("max_phon+adam", 64, 15, 0, 0.03, "adam"),
    ("l2+adam", 512, 10, 0, 0.01, "adam"),
]
STFT_NORMS = ("min_max_freqs", "max_phon", "fletcher_munson")


def case_norm(name):
    return name.split("+")[0].split("_universal")[0].split("_128")[0]


def sweep_bytes(norm, opt, rows, B, T, exact=False):
    """Algorithmic bytes per call."""
    step = 28 if opt == "adam" else 12
    n, m = rows * T, B * T
    if norm == "linf":
        return step * n
    if norm in ("l2", "snr", "tv"):
        return (step + 8) * n + (4 * m if norm != "l2" else 0)
    if norm == "fletcher_munson":
        return (step + 8) * n
    # min_max_freqs / max_phon: fully fused (PGD) -- Adam runs as a streaming pre-pass (28) + the fused transform (8)
    return (step if opt == "pgd" else step + 8) * n


def sweep_flops(norm, rows, T, n_fft=1024, hop=256, exact=False):
    """Algorithmic fp32 flops of the STFT family: 2.5 n log2 n per real transform and frame,
    + window / overlap-add / envelope (~6 n per frame for a round trip, 2 n one way), + the per-bin operator
    (mask 2, phon clip 12, fletcher_munson weight 25 flops per bin)."""
    frames = rows * (1 + T // hop)
    F = n_fft // 2 + 1
    fft = 2.5 * n_fft * math.log2(n_fft)
    if norm == "fletcher_munson":
        f = fft + 2 * n_fft + 25 * F
        if exact:
            f += 2 * fft + 6 * n_fft + 2 * F
        return frames * f
    return frames * (2 * fft + 6 * n_fft + (2 if norm == "min_max_freqs" else 12) * F)


def sweep_inputs(dev, B, sec, sigma, rows=0):
    T = sec * SR
    rows = rows or B
    g = torch.Generator(device=dev).manual_seed(1234)
    clean = (torch.rand(B, T, generator=g, device=dev) * 2 - 1) * 0.1
    p = torch.randn(rows, T, generator=g, device=dev) * sigma
    grad = torch.randn(rows, T, generator=g, device=dev)
    return clean, p, grad


def sweep_args(name, opt, dev):
    from paa_b200.training_utils import parser as pparser
    args = pparser.create_arg_parser().parse_args(["--norm_type", case_norm(name), "--optimizer_type", opt, "--snr_db", "40"])
    args.device = str(dev)
    args.fm_exact_roundtrip = name.endswith("+exact")       # the reference's literal second transform instead of s*q
    return args


def sustained_us(calls, reps: int) -> float:
    """Mean device time per call of `calls` run back to back `reps` times inside ONE event pair (CUDA events tick every
    2.048 us on this box: too coarse for one 10-30 us call).  The callers pass calls on ROTATING input sets that together
    exceed the 126 MB L2 several times, so every call streams cold data; the GPU is kept busy while the host enqueues."""
    for c in calls:
        c()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(10_000_000)
    e0.record()
    for _ in range(reps):
        for c in calls:
            c()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * len(calls))


SUSTAINED_FOOTPRINT = 600e6          # bytes the rotating input sets of `sustained_us` cover together


def projection_sweep(dev, iters: int = 20, only=None):
    """Step + projection alone at BASELINE.json's shapes.  `ms`: CUDA events around ONE call behind a busy GPU (no launch
    latency in the interval), L2 flushed by a 256 MB memset before every call; the figure includes the kernel's ramp-up
    and tail on an otherwise idle GPU and is quantised by the 2.048 us event tick.  `sustained_us` (+ `sustained_GB/s`,
    `sustained_frac`): the same call back to back on rotating cold input sets inside one event pair (sustained_us).
    `wall_ms`: host wall clock of the same call from an idle GPU to a synchronised result -- measured exactly like
    `torch_eager_ms` (baseline_torch_eager)."""
    import paa_b200
    from paa_b200.core import iso
    from paa_b200.training_utils import build as pbuild
    peak, _ = measured_peak()
    interp = iso.build_weight_interpolator()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    out = {}
    for name, B, sec, rows, sigma, opt in SWEEP_CASES:
        if only and name not in only:
            continue
        norm = case_norm(name)
        T = sec * SR
        clean, p, grad = sweep_inputs(dev, B, sec, sigma, rows)
        nrows = p.shape[0]
        args = sweep_args(name, opt, dev)
        thr = pbuild.init_phon_threshold_tensor(args)
        optim = None
        if opt == "adam":
            optim, _ = pbuild.create_optimizer(args, p)
        for _ in range(3):
            paa_b200.step_and_project(p, grad, clean, args, interp, thr, optimizer=optim)
        times = []
        for _ in range(iters):
            flush.zero_()
            torch.cuda._sleep(400_000)            # keep the GPU busy while the host enqueues: no launch latency in the interval
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            paa_b200.step_and_project(p, grad, clean, args, interp, thr, optimizer=optim)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        walls = []
        for _ in range(iters):
            torch.cuda.synchronize()
            w0 = time.perf_counter()
            paa_b200.step_and_project(p, grad, clean, args, interp, thr, optimizer=optim)
            torch.cuda.synchronize()
            walls.append((time.perf_counter() - w0) * 1e3)
        ms = statistics.median(times)
        nbytes = sweep_bytes(norm, opt, nrows, B, T)
        # sustained: rotate over input sets (the first is the one above)
        nsets = max(2, min(24, math.ceil(SUSTAINED_FOOTPRINT / max(nbytes, 1))))
        sets = [(clean, p, grad, optim)]
        for _ in range(nsets - 1):
            c2, p2, g2 = sweep_inputs(dev, B, sec, sigma, rows)
            sets.append((c2, p2, g2, pbuild.create_optimizer(args, p2)[0] if opt == "adam" else None))
        sus = sustained_us([(lambda q=q: paa_b200.step_and_project(q[1], q[2], q[0], args, interp, thr, optimizer=q[3])) for q in sets],
                           max(2, 60 // nsets))
        del sets
        gbs = nbytes / (ms * 1e-3) / 1e9
        rec = {"shape": f"{B}x{sec}s", "p_rows": nrows, "optimizer": opt, "algorithmic_bytes": nbytes, "ms": round(ms, 4),
               "wall_ms": round(statistics.median(walls), 4), "GB/s": round(gbs, 1),
               "frac_of_measured_peak": round(gbs / peak, 4), "audio_s_per_s": round(B * sec / (ms * 1e-3), 1),
               "sustained_us": round(sus, 2), "sustained_GB/s": round(nbytes / sus / 1e3, 1),
               "sustained_frac": round(nbytes / sus / 1e3 / peak, 4)}
        if norm in STFT_NORMS:
            fl = sweep_flops(norm, nrows, T, exact=name.endswith("+exact"))
            rec["gflops"] = round(fl / (ms * 1e-3) / 1e9, 1)
            rec["frac_of_fp32_peak"] = round(fl / (ms * 1e-3) / 1e12 / FP32_PEAK_TFLOPS, 4)
            rec["sustained_frac_of_fp32_peak"] = round(fl / (sus * 1e-6) / 1e12 / FP32_PEAK_TFLOPS, 4)
        out[name] = rec
        del clean, p, grad, optim
        torch.cuda.empty_cache()
    if not only or "compose" in only:
        out.update(compose_sweep(dev, flush, peak, iters))
    return out


def compose_sweep(dev, flush, peak, iters):
    """The input side: x_adv = clamp(clean + p) and dL/dp, universal (1,T) and per-utterance p,
    batch 32 x 10 s and 128 x 10 s; torch eager (add, clamp_, autograd's mask + batch sum) timed beside it."""
    from paa_b200.core.compose import compose_clamp
    res = {}
    for B in (32, 128):
        T = SECONDS * SR
        g = torch.Generator(device=dev).manual_seed(7)
        clean = (torch.rand(B, T, generator=g, device=dev) * 2 - 1) * 0.9
        w = torch.randn(B, T, generator=g, device=dev)
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
            # sustained: back to back on rotating cold input sets (sustained_us)
            nsets = max(2, min(24, math.ceil(SUSTAINED_FOOTPRINT / fb)))
            sets = [(clean, w, p)]
            for _ in range(nsets - 1):
                sets.append(((torch.rand(B, T, generator=g, device=dev) * 2 - 1) * 0.9, torch.randn(B, T, generator=g, device=dev),
                             (torch.randn(rows, T, generator=g, device=dev) * 0.3).requires_grad_(True)))
            f_us = sustained_us([(lambda q=q: compose_clamp(q[0], q[2])) for q in sets], max(2, 60 // nsets))
            xs = [compose_clamp(q[0], q[2]) for q in sets]
            b_us = sustained_us([(lambda q=q, y=y: torch.autograd.grad(y, q[2], q[1], retain_graph=True)) for q, y in zip(sets, xs)],
                                max(2, 60 // nsets))
            del sets, xs
            tag = ("universal" if rows == 1 else "per_utterance") + ("" if B == 32 else f"_{B}x10s")
            res[f"compose_fwd_{tag}"] = {"shape": f"{B}x{SECONDS}s", "ms": round(fwd_ms, 4), "GB/s": round(fb / fwd_ms / 1e6, 1),
                                         "frac_of_measured_peak": round(fb / fwd_ms / 1e6 / peak, 4), "torch_eager_ms": round(tf_ms, 4),
                                         "sustained_us": round(f_us, 2), "sustained_frac": round(fb / f_us / 1e3 / peak, 4)}
            res[f"compose_bwd_{tag}"] = {"shape": f"{B}x{SECONDS}s", "ms": round(bwd_ms, 4), "GB/s": round(bb / bwd_ms / 1e6, 1),
                                         "frac_of_measured_peak": round(bb / bwd_ms / 1e6 / peak, 4), "torch_eager_ms": round(tb_ms, 4),
                                         "sustained_us": round(b_us, 2), "sustained_frac": round(bb / b_us / 1e3 / peak, 4)}
            del p, x, xt
        del clean, w
        torch.cuda.empty_cache()
    return res


# ---------------------------------------------------------------------------------------------------------
REF_SAMPLE = 2           # utterances per step of the CPU legs: a bounded sample of the 512-utterance batch


def oracle_step_seconds(batch: int, steps: int, warmup: int, rows: int):
    """The CPU oracle port of the reference's attack iteration (configs[4]: untargeted l2, PGD) on `batch` x 10 s;
    returns seconds per step."""
    from oracle import paa_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    from transformers import Wav2Vec2Config, Wav2Vec2ForCTC
    torch.manual_seed(0)
    model = Wav2Vec2ForCTC(Wav2Vec2Config()).eval()       # parameters keep requires_grad=True, as the reference leaves them
    T = SECONDS * SR
    clean, p = synth(0, batch, T, "cpu", rows)
    hp = orc.Hyper(norm_type=NORM, optimizer_type="pgd", attack_mode="untargeted", lr=LR)
    p = orc.constrain(p, clean, hp)
    for _ in range(warmup):
        p, _, _ = orc.attack_iteration(model, p, clean, [UNTARGETED_TEXT] * batch, hp)
    t0 = time.perf_counter()
    for _ in range(steps):
        p, loss, _ = orc.attack_iteration(model, p, clean, [UNTARGETED_TEXT] * batch, hp)
    return (time.perf_counter() - t0) / steps


def baseline_torch_eager(dev, sweep):
    """Part of the baseline leg (the only place the oracle port runs): the same torch ops the reference executes,
    eager, on THIS GPU -- the kernel-for-kernel bar of SURVEY.md 2.2 -- per sweep entry at the same shapes and seeds,
    written next to our numbers together with the parity of the two results (PGD and Adam entries alike).
    fletcher_munson includes the reference's D2H -> host bilinear interpolation -> H2D round trip."""
    import paa_b200
    from oracle import paa_oracle as orc
    from paa_b200.core import iso
    from paa_b200.training_utils import build as pbuild
    interp = iso.build_weight_interpolator()
    it_cpu = orc.build_weight_interpolator()
    for name, B, sec, rows, sigma, opt in SWEEP_CASES:
        if name not in sweep:
            continue
        norm = case_norm(name)
        try:
            clean, p, grad = sweep_inputs(dev, B, sec, sigma, rows)
            args = sweep_args(name, opt, dev)
            thr = pbuild.init_phon_threshold_tensor(args)
            hp = orc.Hyper(norm_type=norm, optimizer_type=opt, snr_db=40.0)
            reps = 2 if norm == "fletcher_munson" else 5

            def ref_once():
                st = orc.AdamState(m=torch.zeros_like(p), v=torch.zeros_like(p)) if opt == "adam" else None
                return orc.step_and_constrain(p, grad, clean, hp, it_cpu, thr, adam=st)

            want = ref_once()
            walls = []
            for _ in range(reps):
                torch.cuda.synchronize()
                w0 = time.perf_counter()
                orc.step_and_constrain(p, grad, clean, hp, it_cpu, thr,
                                       adam=orc.AdamState(m=torch.zeros_like(p), v=torch.zeros_like(p)) if opt == "adam" else None)
                torch.cuda.synchronize()
                walls.append((time.perf_counter() - w0) * 1e3)
            eager_ms = statistics.median(walls)
            optim = pbuild.create_optimizer(args, p)[0] if opt == "adam" else None
            got = paa_b200.step_and_project(p, grad, clean, args, interp, thr, optimizer=optim)
            sweep[name]["torch_eager_ms"] = round(eager_ms, 3)
            sweep[name]["speedup_vs_torch_eager"] = round(eager_ms / sweep[name]["wall_ms"], 1)      # wall clock against wall clock
            sweep[name]["max_rel_err_vs_torch_eager"] = float(f"{float((got - want).abs().max() / want.abs().max()):.2e}")
            del want, got, clean, p, grad
        except Exception as exc:                                  # the comparator must never break the bench line
            sweep[name]["torch_eager_ms"] = f"unavailable: {type(exc).__name__}"
        torch.cuda.empty_cache()


def transcript_identity(model, dev):
    """north_star: identical post-attack transcript and WER.  oracle/transcript_check.py runs the reference arithmetic
    (the oracle port on CUDA tensors) and libpaa side by side on wav2vec2-base for 20 steps: configs[1] (targeted snr 40 dB,
    32 x 10 s) teacher-forced and free-running, and configs[2] (max_phon, 64 x 15 s) teacher-forced."""
    try:
        from oracle import paa_oracle as orc, transcript_check as tc
        from paa_b200.core import iso
        from paa_b200.training_utils import build as pbuild
        out = {"margin": 5e-3, "margin_free_running": 2e-2, "margin_fp32_gradient_source": 2e-5,
               "note": "a flip = a logit frame whose greedy token differs; margin = the reference's top-1 minus top-2 logit "
                       "on that frame (logit std ~0.5); flips_above_margin must be false.  control_one_ulp = the reference "
                       "against itself with a random half of the samples of p moved by one fp32 ulp: the gradient source's "
                       "TF32 convolutions (torch default) turn last-bit differences of p into ~1e-3 logit differences; "
                       "with cudnn_tf32 off they stay ~1e-6 (configs1_snr_fp32_gradient_source)"}
        for tag, norm, mode, B, sec, free, micro, tf32, margin in (
                ("configs1_snr", "snr", "targeted", 32, 10, True, 0, None, 5e-3),
                ("configs1_snr_fp32_gradient_source", "snr", "targeted", 32, 10, False, 0, False, 2e-5),
                ("configs2_max_phon", "max_phon", "untargeted", 64, 15, False, 32, None, 5e-3)):
            T = sec * SR
            g = torch.Generator().manual_seed(1234)
            clean = ((torch.rand(B, T, generator=g) * 2 - 1) * 0.1).to(dev)
            p0 = (torch.randn(1, T, generator=g) * 0.01).to(dev)
            args = make_args(norm, mode, dev, snr_db=40.0)
            hp = orc.Hyper(norm_type=norm, optimizer_type="pgd", attack_mode=mode, lr=LR, snr_db=40.0, device=str(dev))
            thr = pbuild.init_phon_threshold_tensor(args)
            rep = tc.run(model, clean, [UNTARGETED_TEXT] * B, args, hp, 20, p0, orc.build_weight_interpolator(),
                         iso.build_weight_interpolator(), thr, micro=micro, free_running=free, cudnn_tf32=tf32)
            for k in ("teacher_forced", "free_running", "control_one_ulp"):
                if k in rep:
                    rep[k]["flips_above_margin"] = bool(tc.flips_above(rep[k], out["margin_free_running"] if k == "free_running" else margin))
            out[tag] = rep
            del clean, p0
            torch.cuda.empty_cache()
        return out
    except Exception as exc:                                      # the checker must never break the bench line
        return {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}


def cpu_baseline(a):
    sec = oracle_step_seconds(REF_SAMPLE, steps=3, warmup=1, rows=1 if a.universal else REF_SAMPLE)
    return {"value": round(REF_SAMPLE * SECONDS / sec, 3), "unit": "audio-s/s", "cores": torch.get_num_threads(),
            "kind": "port", "sample": f"oracle/paa_oracle.py attack_iteration (untargeted l2, PGD), batch {REF_SAMPLE} x {SECONDS} s "
                                      f"(1/{TOTAL_BATCH // REF_SAMPLE} of the step's batch), 1 warm-up + 3 timed steps"}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows = 1 if a.universal else REF_SAMPLE
    steps = max(a.steps, 3)               # at least three timed steps of the bounded sample
    sec = oracle_step_seconds(REF_SAMPLE, steps=steps, warmup=a.warmup, rows=rows)
    value = REF_SAMPLE * SECONDS / sec
    world = int(os.environ.get("WORLD_SIZE", "1"))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": "audio-s/s", "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(sec * 1e3, 2), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD + ", one perturbation row per utterance; CPU arm: each timed step is a bounded sample "
                               f"of {REF_SAMPLE} of the 512 utterances (audio-s/s is per utterance-second, so the sample "
                               "measures the same quantity)",
                   "total_batch": TOTAL_BATCH, "sample_batch": REF_SAMPLE, "seconds": SECONDS, "p_rows": rows,
                   "norm_type": NORM, "optimizer": "pgd", "timed_steps": steps},
        "cpu_baseline": {"value": round(value, 3), "unit": "audio-s/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"oracle/paa_oracle.py attack_iteration (untargeted l2, PGD), batch {REF_SAMPLE} x {SECONDS} s per step, "
                                   f"{steps} timed steps"},
        "e2e": {"value": round(value, 3), "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--micro-batch", dest="micro_batch", type=int, default=64,
                    help="utterances per wav2vec2 call (the step + projection always sees the whole shard)")
    ap.add_argument("--universal", action="store_true", help="one (1,T) perturbation shared by the batch, as the reference's loop")
    ap.add_argument("--no-cpu", dest="no_cpu", action="store_true", help="skip the cpu_baseline leg (oracle port, torch eager, transcript identity)")
    ap.add_argument("--no-sweep", dest="no_sweep", action="store_true", help="skip the per-norm projection sweep")
    ap.add_argument("--no-secondary", dest="no_secondary", action="store_true", help="skip the configs[1] and universal-p legs")
    ap.add_argument("--no-mode-u", dest="no_mode_u", action="store_true", help="N>1: skip the mode U record")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
