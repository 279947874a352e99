"""Init-time helpers the hot path depends on, with the reference's names
(src/training_utils/build.py:288-359)."""
import torch
from torch.nn import Parameter

try:
    from .. import paa_lib as L
    from .train import perturbation_constraint
except ImportError:
    import paa_lib as L
    from training_utils.train import perturbation_constraint


def init_phon_threshold_tensor(args):
    """Per-bin SPL threshold of the ``max_phon_level`` contour, (1, F, 1) fp32 on args.device
    (build.py:325-348): ISO226(phon)(clip(rfftfreq, 20, 20000))."""
    thr = L.spl_thresh(int(args.n_fft), int(args.sr), float(args.max_phon_level))
    return torch.tensor(thr, dtype=torch.float32, device=args.device).view(1, -1, 1)


def init_perturbation(args, length, spl_thresh, interp, first_batch_data):
    """randn(1, length) projected once onto the constraint set (build.py:288-321); resume-from-checkpoint
    keeps the reference's rule (a .pt file holding the (1, L) fp32 tensor)."""
    import os
    ckpt_path = getattr(args, "resume_from", None)
    if ckpt_path and os.path.isfile(ckpt_path):
        p = torch.load(ckpt_path, map_location=args.device).detach().to(args.device)
    else:
        p = torch.randn(1, length, device=args.device)
        p = perturbation_constraint(p=p, clean_audio=first_batch_data, args=args, interp=interp,
                                    spl_thresh=spl_thresh).detach()
    if args.optimizer_type == "adam":
        p = Parameter(p)
    elif args.optimizer_type == "pgd":
        p.requires_grad_()
        p.retain_grad()
    else:
        raise NotImplementedError(f"Unsupported optimizer_type: {args.optimizer_type}")
    if p.shape[-1] != length:
        raise ValueError(f"Loaded perturbation length {p.shape[-1]} != expected {length}")
    return p


class PerturbationAdam(torch.optim.Adam):
    """torch.optim.Adam whose update runs in libpaa.so.  State layout, param_groups and state_dict are
    torch's own (so StepLR and checkpoints behave as in build.py:352-359); ``step()`` launches the
    stand-alone Adam kernel, and ``fused_step_descriptor`` hands the state to a projection kernel so
    that step and projection share one pass over HBM."""

    def _state_for(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.tensor(0.0, dtype=torch.float32)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def _group_of(self, p):
        for g in self.param_groups:
            if any(q is p for q in g["params"]):
                return g
        raise KeyError("parameter is not managed by this optimizer")

    def fused_step_descriptor(self, p, grad):
        """Describe the NEXT update (step count state['step'] + 1) for the C ABI (struct paa_step).  The counter itself
        advances in ``commit_step`` once the launch that consumed the descriptor has been accepted, so a call that
        raises (bad shape, CUDA error) leaves the optimiser state where it was."""
        key = next(q for g in self.param_groups for q in g["params"] if q.data_ptr() == p.data_ptr() or q is p)
        g = self._group_of(key)
        if g["weight_decay"] != 0 or g["amsgrad"] or g["maximize"]:
            raise NotImplementedError("PerturbationAdam supports the reference's configuration only "
                                      "(no weight decay / amsgrad / maximize)")
        st = self._state_for(key)
        desc = L.make_step(L.STEP_ADAM, grad, g["lr"], st["exp_avg"], st["exp_avg_sq"], int(st["step"].item()) + 1,
                           g["betas"], g["eps"])
        desc._adam_state = st
        return desc

    @staticmethod
    def commit_step(desc):
        """The update described by ``desc`` was enqueued: state['step'] += 1 (torch/optim/adam.py:457)."""
        desc._adam_state["step"] += 1

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                L.need_cuda(p)
                desc = self.fused_step_descriptor(p, L.f32c(p.grad))
                plan = L.plan_plain(p)
                rows, T = p.numel() // p.shape[-1], p.shape[-1]
                L.check(L.lib.paa_step_only(plan.h, p.data_ptr(), p.data_ptr(), rows, T, L.step_ref(desc),
                                            L.stream_ptr(p.device)), plan.h)
                self.commit_step(desc)
        return loss


def create_optimizer(args, p):
    """Adam on the perturbation + StepLR(step_size, gamma) (build.py:352-359)."""
    optimizer = PerturbationAdam([p], lr=args.lr)
    scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=args.step_size, gamma=args.gamma)
    return optimizer, scheduler
