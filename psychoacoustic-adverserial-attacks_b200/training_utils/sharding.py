"""Utterance sharding across the GPUs of one box (mode R): every rank attacks a
contiguous block of utterances with its own perturbation rows, optimiser state and projection scalars, so the hot
path needs no collective.  The only exchange of a run is the sum of the WER counters."""
from typing import Sequence, Tuple

import torch
import torch.distributed as dist

try:
    from .. import paa_lib as L
except ImportError:
    import paa_lib as L


def shard_bounds(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows [lo, hi) owned by `rank`: contiguous blocks, sizes differing by at most one, earlier ranks larger."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_rows(batch: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_bounds(batch.shape[0], rank, world)
    return batch[lo:hi]


def allreduce_wer(references: Sequence[str], hypotheses: Sequence[str], device=None) -> Tuple[int, int, float]:
    """(edit errors, reference words, WER) over all ranks: local counters from libpaa's paa_wer_counts, then ONE
    all-reduce(sum) of an int64[2] (NCCL on GPUs, gloo in CPU tests).  Without an initialised process group the
    local counters are returned."""
    e, w = L.wer_counts(list(references), list(hypotheses))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([e, w], dtype=torch.int64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        e, w = int(t[0]), int(t[1])
    return e, w, (e / w if w else 0.0)
