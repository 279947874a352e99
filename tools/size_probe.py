import sys, statistics, torch
sys.path.insert(0, "/root/repo")
import paa_b200
from paa_b200.training_utils import parser as pparser
dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
def run(norm, B, sec, step=True, cold=True):
    T = sec * 16000
    g = torch.Generator(device=dev).manual_seed(1)
    clean = (torch.rand(B, T, generator=g, device=dev) * 2 - 1) * 0.1
    p = torch.randn(B, T, generator=g, device=dev) * 0.01
    grad = torch.randn(B, T, generator=g, device=dev)
    args = pparser.create_arg_parser().parse_args(["--norm_type", norm, "--optimizer_type", "pgd", "--snr_db", "40"]); args.device = "cuda:0"
    ts = []
    for _ in range(12):
        if cold: flush.zero_()
        torch.cuda._sleep(400000)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        if step: paa_b200.step_and_project(p, grad, clean, args, None, None)
        else: paa_b200.perturbation_constraint(p, clean, args, None, None)
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)
for norm in ("l2", "snr", "tv"):
    for B in (32, 128, 512):
        bpe = 20 if norm == "l2" else 24
        us = run(norm, B, 10)
        usp = run(norm, B, 10, step=False)
        n = B * 160000
        print(f"{norm:4s} {B:4d}x10s  pgd {us:8.1f} us  {bpe*n/us/1e3:7.0f} GB/s   proj-only {usp:8.1f} us {((bpe-8)*n)/usp/1e3:7.0f} GB/s", flush=True)
