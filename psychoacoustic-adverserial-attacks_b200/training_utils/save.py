"""Checkpoint formats either side of the hot path, kept byte-compatible with the reference so a
run started there can resume here and vice versa: ``perturbation.pt`` is ``torch.save`` of the CPU fp32 (1, L)
tensor (src/training_utils/save.py:155-156, read back by build.py:294-299), ``results.json`` carries the epoch
(save.py:226-256).  Plots and wav export are out of scope."""
import json
import os

import torch


def save_pert(p, path):
    """torch.save(p.detach().cpu(), path) -- save.py:155-156."""
    torch.save(p.detach().cpu(), path)


def load_pert(path, device):
    """What build.init_perturbation does on resume (build.py:294-299)."""
    return torch.load(path, map_location=device).detach().to(device)


def save_json_results(save_dir, norm_type, attack_size, **kwargs):
    """results.json with the reference's field rules (save.py:226-256): None fields dropped, dict values rounded
    to 4 decimals, perturbation_efficiency = perturbed / clean when both are present."""
    def as_float(v):
        return {k: round(float(v[k]), 4) for k in v} if isinstance(v, dict) else float(v)

    results = {"norm_type": norm_type, "attack_size": float(attack_size)}
    for key, val in kwargs.items():
        if val is not None:
            results[key] = as_float(val)
    clean = kwargs.get("final_test_clean") or kwargs.get("test_loss_clean")
    pert = kwargs.get("final_test_perturbed") or kwargs.get("test_loss_perturbed")
    if clean is not None and pert is not None:
        results["perturbation_efficiency"] = ({k: pert[k] / clean[k] for k in clean} if isinstance(clean, dict)
                                              else pert / clean)
    with open(os.path.join(save_dir, "results.json"), "w") as fh:
        json.dump(results, fh, indent=2)
    return results
