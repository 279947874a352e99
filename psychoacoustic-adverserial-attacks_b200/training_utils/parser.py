"""The hot-path subset of the reference's command line (src/training_utils/parser.py:6-68), same flag
names and defaults, plus ``--device`` which the reference reads but never defines."""
import argparse


def create_arg_parser():
    p = argparse.ArgumentParser()
    p.add_argument("--batch_size", type=int, default=64)
    p.add_argument("--lr", type=float, default=1e-4)
    p.add_argument("--optimizer_type", type=str, choices=["adam", "pgd"], default="adam")
    p.add_argument("--gamma", type=float, default=0.9)
    p.add_argument("--step_size", type=int, default=2)
    p.add_argument("--target_reps", type=int, default=5)
    p.add_argument("--target", type=str, default="delete")
    p.add_argument("--attack_mode", type=str, choices=["untargeted", "targeted"], default="untargeted")
    p.add_argument("--norm_type", type=str, default="max_phon",
                   choices=["l2", "linf", "snr", "tv", "fletcher_munson", "min_max_freqs", "max_phon"])
    p.add_argument("--fm_epsilon", type=float, default=2)
    p.add_argument("--l2_size", type=float, default=0.05)
    p.add_argument("--linf_size", type=float, default=0.0001)
    p.add_argument("--snr_db", type=float, default=64)
    p.add_argument("--min_freq_attack", type=float, default=120)
    p.add_argument("--max_freq_attack", type=float, default=20_000)
    p.add_argument("--tv_epsilon", type=float, default=0.001)
    p.add_argument("--max_phon_level", type=float, default=20)
    p.add_argument("--phon_reference_db", type=float, default=65)
    p.add_argument("--sr", type=int, default=16000)
    p.add_argument("--n_fft", type=int, default=1024)
    p.add_argument("--hop_length", type=int, default=256)
    p.add_argument("--win_length", type=int, default=1024)
    p.add_argument("--seed", type=int, default=5)
    p.add_argument("--device", type=str, default="cuda")
    # not in the reference: fletcher_munson's second pass as the literal ISTFT(scale * STFT(q)) of projections.py:116-133
    # instead of the algebraically identical scale * q on the reconstructed span (the default, ~5e-7 apart)
    p.add_argument("--fm_exact_roundtrip", action="store_true")
    # not in the reference: run the model call of train.py:136-145 in chunks of this many utterances (0 = whole batch).
    # The CTC reduction is "sum", so chunk losses and chunk gradients add up to the whole-batch ones; the step and the
    # projection always see the whole batch.
    p.add_argument("--micro_batch", type=int, default=0)
    # not in the reference: x_adv = clamp(clean + p) and its backward as libpaa kernels instead of autograd's passes
    p.add_argument("--fused_compose", action="store_true")
    # not in the reference: keep loss / greedy ids on the device and decode them after the epoch (no per-step host sync)
    p.add_argument("--defer_metrics", action="store_true")
    return p
