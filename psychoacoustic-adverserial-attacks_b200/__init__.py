"""B200-native (sm_100a) drop-in for the perturbation hot path of
tomer-erez/Psychoacoustic-adverserial-attacks: the PGD/Adam step on the waveform perturbation
fused with the ``norm_type`` projection (reference: src/training_utils/train.py:27-99,155-177,
src/core/{projections,fourier_transforms,iso}.py).

Layout mirrors the reference's ``src/`` tree so its call sites keep working:

    paa_b200.training_utils.train.perturbation_constraint(p, clean_audio, args, interp, spl_thresh)
    paa_b200.core.projections.project_{snr,linf,l2,tv,min_max_freqs,fm_norm,phon_level}
    paa_b200.core.fourier_transforms.compute_{stft,istft}
    paa_b200.core.iso.{ISO226, build_weight_interpolator}
    paa_b200.training_utils.build.{init_phon_threshold_tensor, init_perturbation, create_optimizer}

Everything runs in ``libpaa.so`` (hand-written CUDA behind the C ABI of include/paa.h), reached
through ctypes with raw device pointers.  There is no CPU or PyTorch fallback: CUDA tensors or
an exception.
"""
from . import paa_lib                    # noqa: F401  (loads libpaa.so, raises if it is missing)
from . import core, training_utils       # noqa: F401
from .training_utils.train import perturbation_constraint, step_and_project, train_epoch  # noqa: F401

__all__ = ["paa_lib", "core", "training_utils", "perturbation_constraint", "step_and_project", "train_epoch"]
