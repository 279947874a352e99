import sys, time, torch
sys.path.insert(0, "/root/repo")
import paa_b200
from paa_b200.training_utils import parser as pparser, build as pbuild
from paa_b200.core import iso
dev = torch.device("cuda:0")
interp = iso.build_weight_interpolator()
for norm in ("linf", "snr", "max_phon", "fletcher_munson"):
    args = pparser.create_arg_parser().parse_args(["--norm_type", norm, "--optimizer_type", "pgd", "--snr_db", "40"]); args.device = "cuda:0"
    p = torch.randn(1, 16000, device=dev) * 0.01; g = torch.randn(1, 16000, device=dev); c = torch.rand(2, 16000, device=dev) * 0.1
    thr = pbuild.init_phon_threshold_tensor(args)
    for _ in range(20): paa_b200.step_and_project(p, g, c, args, interp, thr)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = 2000
    for _ in range(n): paa_b200.step_and_project(p, g, c, args, interp, thr)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{norm:16s} host {1e6*(t1-t0)/n:6.1f} us/call   incl. drain {1e6*(t2-t0)/n:6.1f} us/call")
