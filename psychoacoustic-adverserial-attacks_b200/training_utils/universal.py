"""Mode U: ONE universal perturbation shared by all GPUs of the box.

Every rank attacks its own utterance shard with the same (1, T) perturbation, so that G GPUs reproduce a
single-process run of the reference on the union of the shards (the CTC reduction is "sum": the gradient of the whole
batch is the sum of the shard gradients, train.py:158; snr / tv need the clean statistics of the whole batch,
projections.py:17, :58).  That makes one real exchange per step: T floats of gradient and two doubles per rank.

The exchange is fused into the hot-path kernels instead of preceding them: each rank copies its partial gradient (and
its clean statistics, paa_clean_stats) into a symmetric-memory buffer (torch.distributed._symmetric_memory: plumbing --
allocation, rendezvous, device-side barrier), and libpaa's step + projection kernels read all G buffers over NVLink
peer access and add them in rank order while they step (include/paa.h, struct paa_parts).  Every rank computes
bit-identical perturbations, nothing but the partials crosses the links, and there is no separate all-reduce pass.
Buffers are double buffered by step parity, so one barrier per step is enough: a rank can only be two steps ahead of a
peer that still reads a slot after it has passed the barrier in between, which that peer joins after its read.

``backend="nccl"`` is the plain baseline (all-reduce, then the ordinary kernels on the reduced buffers);
``backend="auto"`` takes the peer-memory path when the rendezvous succeeds on every rank and the baseline otherwise."""
import logging
from typing import Optional

import torch
import torch.distributed as dist

try:
    from .. import paa_lib as L
except ImportError:
    import paa_lib as L


def reduce_partials(grad: torch.Tensor, stats: torch.Tensor, group=None):
    """Baseline exchange: in-place all-reduce(sum) of the partial gradient and of the (3,) fp64 clean statistics.
    Device agnostic (NCCL on GPUs, gloo in the CPU tests)."""
    dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return grad, stats


def broadcast_perturbation(p: torch.Tensor, src: int = 0, group=None) -> torch.Tensor:
    """All ranks must start from the same bits."""
    dist.broadcast(p, src=src, group=group)
    return p


class UniversalExchange:
    """Per-step exchange of mode U.  ``publish(grad, clean)`` returns the ``paa_parts`` descriptor that
    ``step_and_project(..., parts=...)`` / ``perturbation_constraint(..., parts=...)`` hand to libpaa."""

    def __init__(self, rows: int, T: int, device: torch.device, group=None, backend: str = "symmetric"):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("mode U needs an initialised process group")
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > L.MAX_PARTS:
            raise ValueError(f"mode U supports up to {L.MAX_PARTS} ranks")
        self.rows, self.T, self.device, self.backend = int(rows), int(T), device, backend
        self.n = self.rows * self.T
        self.n_al = (self.n + 3) // 4 * 4                       # statistics start 16-byte aligned
        self.slot = self.n_al + 8                                # [gradient | 3 doubles: sum x^2, sum |dx|, numel | pad]
        self.step = 0
        self._keep = None
        if backend == "auto":
            # peer-memory exchange when the box supports it (NVLink P2P + symmetric memory), else the NCCL baseline.
            # Every rank must take the same branch: the outcome of the rendezvous is agreed on with one all-reduce.
            try:
                self._init_symmetric()
                ok = 1
            except Exception as exc:                                    # noqa: BLE001
                logging.getLogger("asr_attack").warning("mode U: symmetric memory unavailable (%s); using NCCL", exc)
                ok = 0
            flag = torch.tensor([ok], dtype=torch.int32, device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            backend = self.backend = "symmetric" if int(flag.item()) == 1 else "nccl"
            if backend == "nccl":
                self.buf = torch.zeros(2 * self.slot, dtype=torch.float32, device=device)
                self.hdl, self.ptrs = None, None
        elif backend == "symmetric":
            self._init_symmetric()
        elif backend == "nccl":
            self.buf = torch.zeros(2 * self.slot, dtype=torch.float32, device=device)
            self.hdl, self.ptrs = None, None
        else:
            raise ValueError(f"unknown backend {backend!r}")

    def _init_symmetric(self) -> None:
        import torch.distributed._symmetric_memory as symm
        self.buf = symm.empty(2 * self.slot, dtype=torch.float32, device=self.device)
        self.hdl = symm.rendezvous(self.buf, self.group.group_name)
        self.ptrs = [int(x) for x in self.hdl.buffer_ptrs]
        self.buf.zero_()

    def publish(self, grad: Optional[torch.Tensor], clean: Optional[torch.Tensor], norm_type: Optional[str] = None):
        """Make this rank's partial gradient and clean statistics visible to all ranks for this step.
        Everything is enqueued on the current stream, no host synchronisation and no host-side agreement between the
        ranks: the size of the whole clean batch (project_snr's ``clean.numel()``) travels as a third statistic next to
        the two sums and is added up on the device, so shards may be uneven and change from step to step.
        ``norm_type``: when given, the clean statistics are computed only for the norms that use them (snr, tv)."""
        if norm_type is not None and norm_type not in ("snr", "tv"):
            clean = None
        slot = self.step % 2
        self.step += 1
        off = slot * self.slot
        gview = self.buf[off:off + self.n]
        if grad is not None:
            L.need_cuda(grad)
            gview.copy_(grad.detach().reshape(-1))
        clean_numel = 0
        have_stats = clean is not None
        if have_stats:
            L.need_cuda(clean)
            c = L.f32c(clean.detach())
            c2 = c.reshape(-1, c.shape[-1])
            plan = L.plan_plain(c)
            stats_ptr = self.buf.data_ptr() + (off + self.n_al) * 4
            L.check(L.lib.paa_clean_stats(plan.h, c2.data_ptr(), c2.shape[0], c2.shape[1], stats_ptr, plan.scratch(0, 0),
                                          L.stream_ptr(c.device)), plan.h)
        if self.backend == "symmetric":
            self.hdl.barrier(channel=slot)                          # device side, on the current stream
            gp = [p + off * 4 for p in self.ptrs] if grad is not None else None
            sp = [p + (off + self.n_al) * 4 for p in self.ptrs] if have_stats else None
        else:
            stats = self.buf[off + self.n_al:off + self.n_al + 6].view(torch.float64)
            reduce_partials(gview, stats, self.group)
            gp = [gview.data_ptr()] if grad is not None else None
            sp = [stats.data_ptr()] if have_stats else None
        if gp is None and sp is None:
            return None
        return L.make_parts(gp, sp, clean_numel)          # clean_numel = 0: the kernels sum the parts' third statistic
