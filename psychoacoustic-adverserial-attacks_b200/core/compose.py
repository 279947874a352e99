"""The input side of the hot path: ``perturbed = (clean_audio + p).clamp_(-1, 1)`` of
src/training_utils/train.py:136 as one kernel, with its backward (the clamp mask, and for a universal (1,T)
perturbation the sum over the batch) as another.  torch does this in four passes and materialises the mask."""
import torch

try:
    from .. import paa_lib as L
except ImportError:
    import paa_lib as L


class _ComposeClamp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, clean, p):
        L.need_cuda(clean, p)
        c, q = L.f32c(clean.detach()), L.f32c(p.detach())
        if c.dim() != 2 or q.dim() != 2 or q.shape[1] != c.shape[1] or q.shape[0] not in (1, c.shape[0]):
            raise RuntimeError(f"compose_clamp expects (B,T) audio and a (1,T) or (B,T) perturbation, got "
                               f"{tuple(c.shape)} and {tuple(q.shape)}")
        plan = L.plan_plain(c)
        out = torch.empty_like(c)
        L.check(L.lib.paa_compose_clamp(plan.h, c.data_ptr(), c.shape[0], q.data_ptr(), q.shape[0], c.shape[1],
                                        out.data_ptr(), L.stream_ptr(c.device)), plan.h)
        ctx.save_for_backward(c, q)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        c, q = ctx.saved_tensors
        g = L.f32c(grad_out)
        plan = L.plan_plain(c)
        gp = torch.empty_like(q)
        L.check(L.lib.paa_compose_clamp_backward(plan.h, c.data_ptr(), c.shape[0], q.data_ptr(), q.shape[0], c.shape[1],
                                                 g.data_ptr(), gp.data_ptr(), L.stream_ptr(c.device)), plan.h)
        return None, gp


def compose_clamp(clean_audio, p):
    """clamp(clean_audio + p, -1, 1), differentiable with respect to p (clean_audio gets no gradient, train.py:130)."""
    return _ComposeClamp.apply(clean_audio, p)
