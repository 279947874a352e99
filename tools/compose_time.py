"""Times the compose + clamp kernels (forward / backward, universal and per-utterance p) like bench.py's sweep."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
peak, _ = bench.measured_peak()
print(json.dumps(bench.compose_sweep(dev, flush, peak, 20), indent=1))
