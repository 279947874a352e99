"""The N>1 host logic on CPU: utterance sharding and the single WER-counter all-reduce, gloo, world size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import paa_b200  # noqa: F401
from paa_b200.training_utils import sharding

REFS = ["the cat sat on the mat", "hello world", "delete delete delete", "a b c d e", "one two", "x"]
HYPS = ["the cat sat mat", "hello there world", "delete", "a b c d e", "", "y z"]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharding.shard_bounds(len(REFS), rank, world)
        e, w, wer = sharding.allreduce_wer(REFS[lo:hi], HYPS[lo:hi], device="cpu")
        rows = sharding.shard_rows(torch.arange(10).view(5, 2), rank, world)
        out.put((rank, e, w, wer, rows.tolist()))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    for n in (0, 1, 5, 32, 512, 513):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert [sharding.shard_bounds(512, r, 8) for r in (0, 7)] == [(0, 64), (448, 512)]
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 2, 2)


def test_wer_allreduce_world2():
    from paa_b200 import paa_lib as L
    want = L.wer_counts(REFS, HYPS)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, e, w, wer, rows in got:
        assert (e, w) == want and abs(wer - want[0] / want[1]) < 1e-12
    assert got[0][4] == [[0, 1], [2, 3], [4, 5]] and got[1][4] == [[6, 7], [8, 9]]


def test_without_process_group_is_local():
    e, w, wer = sharding.allreduce_wer(["a b"], ["a"])
    assert (e, w, wer) == (1, 2, 0.5)


def _worker_partials(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from paa_b200.training_utils import universal
        g = torch.full((1, 8), float(rank + 1))
        st = torch.tensor([10.0 * (rank + 1), 0.5], dtype=torch.float64)
        universal.reduce_partials(g, st)
        p = torch.full((1, 4), float(rank))
        universal.broadcast_perturbation(p, src=0)
        out.put((rank, g.tolist(), st.tolist(), p.tolist()))
    finally:
        dist.destroy_process_group()


def test_mode_u_baseline_exchange_world2():
    """Mode U's plain exchange (all-reduce of the partial gradient and the clean statistics, broadcast of p), gloo."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_partials, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, g, st, p in got:
        assert g == [[3.0] * 8] and st == [30.0, 1.0] and p == [[0.0] * 4]
