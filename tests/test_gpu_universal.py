"""Mode U on two GPUs: one universal (1,T) perturbation shared by two ranks that hold disjoint
utterance shards.  The kernels read both ranks' partial gradients / clean statistics from peer memory (symmetric
memory) and add them in rank order while they step; the result must equal the oracle's single-process step on the
union of the shards, and must be bit-identical on both ranks.  Run with `gpurun --gpus 2`; skipped on one GPU."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu
NORMS = [("linf", 1e-3), ("l2", 0.01), ("snr", 0.01), ("tv", 0.01), ("max_phon", 0.03), ("min_max_freqs", 0.01),
         ("fletcher_munson", 0.1)]
B, T, STEPS = 4, 8192, 3


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _inputs(step):
    g = torch.Generator().manual_seed(77 + step)
    clean = (torch.rand(B, T, generator=g) * 2 - 1) * 0.1
    grads = torch.randn(2, 1, T, generator=g)
    grads[:, :, ::97] = 0.0                                   # sign(0) = 0 must survive the summation
    grads[1, :, ::97][:, ::2] = 0.0
    return clean, grads


def _worker(rank, world, port, backend, optimizer, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import paa_b200
        from paa_b200.core import iso
        from paa_b200.training_utils import build, parser, universal
        interp = iso.build_weight_interpolator()
        res = {}
        for norm, sigma in NORMS:
            args = parser.create_arg_parser().parse_args(["--norm_type", norm, "--optimizer_type", optimizer,
                                                          "--snr_db", "40", "--lr", "1e-4"])
            args.device = str(dev)
            thr = build.init_phon_threshold_tensor(args)
            p = (torch.randn(1, T, generator=torch.Generator().manual_seed(5)) * sigma).to(dev)
            opt = None
            if optimizer == "adam":
                p = p.requires_grad_(True)
                opt, _ = build.create_optimizer(args, p)
            exch = universal.UniversalExchange(1, T, dev, backend=backend)
            outs = []
            for step in range(STEPS):
                clean, grads = _inputs(step)
                # uneven shards that change from step to step (3+1, 1+3, 3+1): the size of the whole batch is summed on
                # the device from the ranks' statistics, no host-side agreement
                cut = 3 if step % 2 == 0 else 1
                clean_r = (clean[:cut] if rank == 0 else clean[cut:]).to(dev)
                parts = exch.publish(grads[rank].to(dev), clean_r, norm)
                with torch.no_grad():
                    q = paa_b200.step_and_project(p.data if opt else p, grads[rank].to(dev), clean_r, args, interp, thr,
                                                  optimizer=opt, parts=parts)
                    if opt:
                        p.data = q
                    else:
                        p = q
                outs.append(q.detach().cpu().clone())
            res[norm] = outs
        out.put((rank, res))
    finally:
        dist.destroy_process_group()


def _run(backend, optimizer):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, backend, optimizer, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return got


def _oracle(optimizer):
    from oracle import paa_oracle as orc
    it = orc.build_weight_interpolator()
    want = {}
    for norm, sigma in NORMS:
        hp = orc.Hyper(norm_type=norm, optimizer_type=optimizer, snr_db=40.0, lr=1e-4)
        thr = orc.phon_threshold(hp.n_fft, hp.sr, hp.max_phon_level)
        p = torch.randn(1, T, generator=torch.Generator().manual_seed(5)) * sigma
        adam = orc.AdamState(m=torch.zeros(1, T), v=torch.zeros(1, T)) if optimizer == "adam" else None
        outs = []
        for step in range(STEPS):
            clean, grads = _inputs(step)
            g = grads[0] + grads[1]                             # rank order, fp32: what the kernels add
            p = orc.step_and_constrain(p, g, clean, hp, it, thr, adam=adam)
            outs.append(p.clone())
        want[norm] = outs
    return want


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="mode U needs two GPUs")
@pytest.mark.parametrize("backend,optimizer", [("symmetric", "pgd"), ("symmetric", "adam"), ("nccl", "pgd"), ("auto", "pgd")])
def test_universal_two_ranks_match_single_process_oracle(backend, optimizer):
    from conftest import rel_max
    got = _run(backend, optimizer)
    want = _oracle(optimizer)
    for norm, _ in NORMS:
        for step in range(STEPS):
            a, b = got[0][norm][step], got[1][norm][step]
            assert torch.equal(a, b), f"{norm} step {step}: ranks disagree"
            err = rel_max(a, want[norm][step])
            assert err <= 1e-5, (backend, optimizer, norm, step, err)


# ---- the attack loop in mode U: train_epoch on two ranks against train_epoch on one rank with the union batch ----
def _loop_worker(rank, world, port, norm, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import paa_b200  # noqa: F401
        from paa_b200.core import loss_helpers
        from paa_b200.training_utils import parser, train, universal
        from test_gpu_attack_loop import tiny_model
        args = parser.create_arg_parser().parse_args(["--norm_type", norm, "--optimizer_type", "pgd", "--snr_db", "30",
                                                      "--lr", "1e-3", "--linf_size", "0.002"])
        args.device = str(dev)
        model = tiny_model(dev)
        for q in model.parameters():
            q.requires_grad_(False)
        g = torch.Generator().manual_seed(3)
        Tl, Bl = 16000, 4
        batches = [(torch.rand(Bl, Tl, generator=g) * 2 - 1) * 0.1 for _ in range(2)]
        texts = ["hello world this is a test"] * Bl
        p0 = (torch.randn(1, Tl, generator=g) * 1e-3).to(dev)
        # two ranks, half of every batch each, one shared perturbation
        lo, hi = rank * Bl // world, (rank + 1) * Bl // world
        args.universal_exchange = universal.UniversalExchange(1, Tl, dev)
        res_u = train.train_epoch(args, [(b[lo:hi], texts[lo:hi]) for b in batches], p0.clone(), model, 0, None, None,
                                  loss_helpers.WerMetric(), None, None)
        pu = res_u.p.detach().cpu()
        # the same epoch in one process on the whole batches (rank 0 only needs it, both run it to stay in step)
        args.universal_exchange = None
        res_1 = train.train_epoch(args, [(b, texts) for b in batches], p0.clone(), model, 0, None, None,
                                  loss_helpers.WerMetric(), None, None)
        out.put((rank, pu, res_1.p.detach().cpu()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="mode U needs two GPUs")
@pytest.mark.parametrize("norm", ["linf", "snr"])
def test_universal_train_epoch_reproduces_the_single_process_epoch(norm):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_loop_worker, args=(r, 2, port, norm, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {r: (pu, p1) for r, pu, p1 in (q.get(timeout=600) for _ in procs)}
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert torch.equal(got[0][0], got[1][0]), "ranks hold different perturbations"
    pu, p1 = got[0]
    # The shard gradients are summed in a different order than autograd sums the whole batch, so elements whose
    # gradient is ~0 may take the other sign (a PGD step of 2*lr apart); everything else must agree.
    differ = ((pu - p1).abs() > 1e-6 * p1.abs().max()).float().mean()
    assert float(differ) < 5e-3, float(differ)


# ---- the same kernels on ONE GPU: paa_parts whose buffers are all local ------------------------------------------
# struct paa_parts only holds device pointers; nothing requires them to be peer memory.  Passing 2, 3 and 8 local
# partial-gradient buffers and per-part clean statistics exercises every kStepParts instantiation (k_step_clamp,
# k_fused, k_reduce), k_sum_parts + staging for the STFT-domain projections, and the device-side sum of the clean
# statistics (incl. the batch size), against the oracle's step on the summed gradient and the whole batch.
@pytest.mark.parametrize("optimizer", ["pgd", "adam"])
@pytest.mark.parametrize("G", [2, 3, 8])
@pytest.mark.parametrize("norm,sigma", NORMS)
def test_local_parts_match_oracle_on_summed_gradient(norm, sigma, G, optimizer):
    import paa_b200
    from conftest import rel_l2, rel_max
    from oracle import paa_oracle as orc
    from paa_b200 import paa_lib as L
    from paa_b200.core import iso
    from paa_b200.training_utils import build, parser

    dev = torch.device("cuda:0")
    Tl, Bl = 20000, 9
    g = torch.Generator().manual_seed(100 * G + len(norm))
    args = parser.create_arg_parser().parse_args(["--norm_type", norm, "--optimizer_type", optimizer, "--snr_db", "40",
                                                  "--lr", "1e-4"])
    args.device = str(dev)
    hp = orc.Hyper(norm_type=norm, optimizer_type=optimizer, snr_db=40.0, lr=1e-4)
    it_cpu, it_gpu = orc.build_weight_interpolator(), iso.build_weight_interpolator()
    thr_c, thr_g = orc.phon_threshold(hp.n_fft, hp.sr, hp.max_phon_level), build.init_phon_threshold_tensor(args)
    p_ref = torch.randn(1, Tl, generator=g) * sigma
    p_new = p_ref.clone().to(dev)
    opt = None
    if optimizer == "adam":
        p_new = p_new.requires_grad_(True)
        opt, _ = build.create_optimizer(args, p_new)
    adam = orc.AdamState(m=torch.zeros(1, Tl), v=torch.zeros(1, Tl)) if optimizer == "adam" else None
    plan = L.plan_plain(p_new)
    for step in range(2):
        clean = (torch.rand(Bl, Tl, generator=g) * 2 - 1) * 0.1
        grads = torch.randn(G, 1, Tl, generator=g)
        grads[:, :, ::89] = 0.0
        total = grads[0].clone()
        for k in range(1, G):
            total = total + grads[k]                            # left to right in fp32: what the kernels add
        p_ref = orc.step_and_constrain(p_ref, total, clean, hp, it_cpu, thr_c, adam=adam)
        # uneven row shards of the clean batch: the first parts take one row each, the last one the rest
        bounds = list(range(G)) + [Bl] if G < Bl else list(range(Bl)) + [Bl]
        shards = [clean[bounds[k]:bounds[k + 1]] for k in range(len(bounds) - 1)]
        while len(shards) < G:                                    # more parts than rows: empty-statistics parts
            shards.append(None)
        g_dev = [grads[k].to(dev).contiguous() for k in range(G)]
        stats = torch.zeros(G, 4, dtype=torch.float64, device=dev)
        keep = []
        for k, sh in enumerate(shards):
            if sh is None:
                continue
            c = sh.to(dev).contiguous()
            keep.append(c)
            L.check(L.lib.paa_clean_stats(plan.h, c.data_ptr(), c.shape[0], c.shape[1], stats[k].data_ptr(),
                                          plan.scratch(0, 0), L.stream_ptr(dev)), plan.h)
        parts = L.make_parts([t.data_ptr() for t in g_dev], [stats[k].data_ptr() for k in range(G)], 0)
        clean_local = keep[0]                                     # a rank only holds its own shard
        with torch.no_grad():
            q = paa_b200.step_and_project(p_new.data if opt else p_new, g_dev[0], clean_local, args, it_gpu, thr_g,
                                          optimizer=opt, parts=parts)
        if opt:
            p_new.data = q
        else:
            p_new = q
        a, b = rel_max(q.cpu(), p_ref), rel_l2(q.cpu(), p_ref)
        assert a <= 1e-5 and b <= 1e-5, (norm, G, optimizer, step, a, b)
