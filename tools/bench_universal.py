"""Mode U exchange + step + projection alone (torchrun, N >= 2 GPUs): the fused peer-memory path against the plain
NCCL all-reduce followed by the ordinary kernels.  Universal (1,T) perturbation, 10 s of audio, batch 32 per rank.
Prints per-norm microseconds per step (CUDA events, max over ranks)."""
import json
import os
import statistics
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import paa_b200  # noqa: E402
from paa_b200.training_utils import build as pbuild, parser as pparser, universal  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, T, ITERS = 32, 160000, 50
g = torch.Generator(device=dev).manual_seed(1234 + rank)
clean = (torch.rand(B, T, generator=g, device=dev) * 2 - 1) * 0.1
grad = torch.randn(1, T, generator=g, device=dev)
out = {}
for norm, sigma in (("linf", 1e-3), ("l2", 0.01), ("snr", 0.01), ("tv", 0.01), ("max_phon", 0.03)):
    args = pparser.create_arg_parser().parse_args(["--norm_type", norm, "--optimizer_type", "pgd", "--snr_db", "40"])
    args.device = str(dev)
    thr = pbuild.init_phon_threshold_tensor(args)
    p = universal.broadcast_perturbation(torch.randn(1, T, device=dev) * sigma)
    for backend in ("symmetric", "nccl"):
        exch = universal.UniversalExchange(1, T, dev, backend=backend)
        times = []
        for it in range(ITERS + 5):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            parts = exch.publish(grad, clean, norm)
            q = paa_b200.step_and_project(p, grad, clean, args, None, thr, parts=parts)
            e1.record()
            torch.cuda.synchronize()
            if it >= 5:
                times.append(e0.elapsed_time(e1))
        t = torch.tensor([statistics.median(times)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[f"{norm}/{backend}"] = round(float(t) * 1e3, 1)
if rank == 0:
    print(json.dumps({"what": "mode U: publish (copy + clean statistics + exchange) + step + projection, us per step, "
                              f"universal (1,{T}) p, clean {B}x{T} per rank, {world} GPUs", "us": out}), flush=True)
dist.destroy_process_group()
