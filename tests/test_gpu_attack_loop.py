"""The hot path driven the way the reference drives it: train_epoch (train.py:103-182) on a small random wav2vec2,
PGD and Adam, every norm_type.  Checks the loop mechanics (state, shapes, constraint after each step, WER counters);
numerical parity of the step + projection itself is in test_gpu_parity.py."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def tiny_model(dev):
    from transformers import Wav2Vec2Config, Wav2Vec2ForCTC
    torch.manual_seed(0)
    cfg = Wav2Vec2Config(hidden_size=64, num_hidden_layers=2, num_attention_heads=2, intermediate_size=128,
                         conv_dim=(32,) * 7, num_conv_pos_embeddings=16, num_conv_pos_embedding_groups=4, vocab_size=32)
    return Wav2Vec2ForCTC(cfg).eval().to(dev)


@pytest.mark.parametrize("opt", ["pgd", "adam"])
@pytest.mark.parametrize("norm", ["linf", "l2", "snr", "tv", "min_max_freqs", "max_phon", "fletcher_munson"])
def test_train_epoch_runs_and_constrains(norm, opt):
    import paa_b200
    from paa_b200.core import iso, loss_helpers
    from paa_b200.training_utils import build, parser, train
    dev = torch.device("cuda:0")
    args = parser.create_arg_parser().parse_args(
        ["--norm_type", norm, "--optimizer_type", opt, "--snr_db", "30", "--lr", "1e-3", "--l2_size", "0.5",
         "--linf_size", "0.002", "--fm_epsilon", "5", "--attack_mode", "targeted" if norm == "snr" else "untargeted"])
    args.device = str(dev)
    args.fused_compose = (opt == "adam")               # exercise both composes: autograd's and libpaa's kernels
    model = tiny_model(dev)
    for q in model.parameters():
        q.requires_grad_(False)
    g = torch.Generator().manual_seed(3)
    T = 16000
    loader = [((torch.rand(3, T, generator=g) * 2 - 1) * 0.1, ["hello world this is a test"] * 3) for _ in range(3)]
    interp = iso.build_weight_interpolator()
    thr = build.init_phon_threshold_tensor(args)
    first = loader[0][0].to(dev)
    p = build.init_perturbation(args, T, thr, interp, first)
    assert p.shape == (1, T) and p.requires_grad
    optimizer = scheduler = None
    if opt == "adam":
        optimizer, scheduler = build.create_optimizer(args, p)
    wer = loss_helpers.WerMetric()
    res = train.train_epoch(args, loader, p, model, 0, None, interp, wer, thr, optimizer)
    p_new, ctc, w = res                                   # tuple-unpacks like run_attack.py:64 expects
    assert p_new.shape == (1, T) and torch.isfinite(p_new).all()
    assert ctc > 0 and 0.0 <= w and wer.words == 3 * 3 * 6
    assert not torch.equal(p_new.detach(), p.detach()) or opt == "adam"
    last = loader[-1][0].to(dev)
    pd = p_new.detach()
    if norm == "linf":
        assert float(pd.abs().max()) <= args.linf_size * (1 + 1e-6)
    elif norm == "l2":
        assert float(pd.norm()) <= args.l2_size * (1 + 1e-5)
    elif norm == "snr":
        # the reference sizes the target norm with clean.numel() even for a (1,T) perturbation (projections.py:29),
        # so against a batch of 3 the per-sample SNR may sit 10*log10(3) dB under snr_db
        snr = 10 * torch.log10(last.pow(2).mean() / (pd.pow(2).mean() + 1e-12))
        assert float(snr) >= args.snr_db - 10 * torch.log10(torch.tensor(3.0)).item() - 1e-3
    elif norm == "tv":
        tv = lambda x: (x[:, 1:] - x[:, :-1]).abs().sum()        # noqa: E731
        assert float(tv(pd)) <= float(args.tv_epsilon * tv(last)) * (1 + 1e-4)
    else:
        assert float(pd[:, 256 * (T // 256):].abs().max() if T % 256 else 0.0) == 0.0
    if opt == "adam":
        st = optimizer.state[p]
        assert int(st["step"]) == 3 and st["exp_avg"].abs().sum() > 0
        scheduler.step()                                    # StepLR keeps working on the drop-in optimiser
        sd = optimizer.state_dict()
        assert sd["state"][0]["exp_avg"].shape == (1, T)


def test_deferred_metrics_equal_synchronous():
    """--defer_metrics changes when the loss and WER are read back, not what they are."""
    import paa_b200  # noqa: F401
    from paa_b200.core import loss_helpers
    from paa_b200.training_utils import build, parser, train
    dev = torch.device("cuda:0")
    model = tiny_model(dev)
    for q in model.parameters():
        q.requires_grad_(False)
    g = torch.Generator().manual_seed(4)
    loader = [((torch.rand(2, 16000, generator=g) * 2 - 1) * 0.1, ["hello world this is a test"] * 2) for _ in range(3)]
    out = []
    for defer in (False, True):
        args = parser.create_arg_parser().parse_args(["--norm_type", "linf", "--optimizer_type", "pgd", "--linf_size", "0.002",
                                                      "--lr", "1e-3"] + (["--defer_metrics"] if defer else []))
        args.device = str(dev)
        torch.manual_seed(1)
        p = build.init_perturbation(args, 16000, None, None, None)
        wer = loss_helpers.WerMetric()
        res = train.train_epoch(args, loader, p, model, 0, None, None, wer, None, None)
        out.append((res.p.detach().clone(), res.avg_ctc, res.avg_wer, wer.errors, wer.words))
    assert torch.equal(out[0][0], out[1][0]) and out[0][1:] == out[1][1:]


@pytest.mark.parametrize("norm", ["snr", "max_phon", "linf"])
def test_cuda_graph_capture(norm):
    """The fused step + projection makes no host synchronisation and allocates nothing itself, so it can be captured
    in a CUDA graph -- including the cooperative launch of the reducing norms."""
    import paa_b200
    from paa_b200.training_utils import build, parser
    dev = torch.device("cuda:0")
    args = parser.create_arg_parser().parse_args(["--norm_type", norm, "--optimizer_type", "pgd", "--snr_db", "40"])
    args.device = str(dev)
    g = torch.Generator(device=dev).manual_seed(2)
    clean = (torch.rand(4, 16000, generator=g, device=dev) * 2 - 1) * 0.1
    p = torch.randn(4, 16000, generator=g, device=dev) * 0.02
    grad = torch.randn(4, 16000, generator=g, device=dev)
    thr = build.init_phon_threshold_tensor(args)
    want = paa_b200.step_and_project(p, grad, clean, args, None, thr)          # also warms handle, scratch, attributes
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = paa_b200.step_and_project(p, grad, clean, args, None, thr)
    for _ in range(3):
        out.zero_()
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, want)
    grad.neg_()                                                                 # replays read the live input buffers
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, paa_b200.step_and_project(p, grad, clean, args, None, thr))


def test_two_streams_share_a_plan_without_sharing_scratch():
    """include/paa.h: entry points are re-entrant as long as concurrent calls use distinct scratch.  The Python binding
    keeps one scratch buffer per CUDA stream, so two streams can drive the same (device, n_fft, hop) plan at once: the
    reducing norms (partials + scalars in scratch) and fletcher_munson (staging buffer + tile partials in scratch) are
    launched back to back on two streams, repeatedly, and must reproduce their single-stream results bit for bit."""
    import paa_b200
    from paa_b200.core import iso
    from paa_b200.training_utils import build, parser
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(5)
    it = iso.build_weight_interpolator()
    jobs = []
    for norm, B, T, sigma in (("l2", 16, 160000, 0.01), ("snr", 8, 160000, 0.01), ("fletcher_munson", 6, 120000, 0.1),
                              ("tv", 12, 80000, 0.01), ("max_phon", 4, 100000, 0.03)):
        args = parser.create_arg_parser().parse_args(["--norm_type", norm, "--optimizer_type", "pgd", "--snr_db", "40"])
        args.device = str(dev)
        clean = (torch.rand(B, T, generator=g, device=dev) * 2 - 1) * 0.1
        p = torch.randn(B, T, generator=g, device=dev) * sigma
        grad = torch.randn(B, T, generator=g, device=dev)
        thr = build.init_phon_threshold_tensor(args)
        want = paa_b200.step_and_project(p, grad, clean, args, it, thr).clone()
        jobs.append((args, clean, p, grad, thr, want))
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    for rep in range(6):
        outs = []
        for k, (args, clean, p, grad, thr, want) in enumerate(jobs):
            with torch.cuda.stream(s1 if (k + rep) % 2 == 0 else s2):
                outs.append(paa_b200.step_and_project(p, grad, clean, args, it, thr))
        torch.cuda.synchronize()
        for (args, *_rest, want), out in zip(jobs, outs):
            assert torch.equal(out, want), (rep, args.norm_type)
