"""Autograd bridge of the training path: `net(inputs, pos_codes)` in training mode returns the reference's
3*num_stages output lists as differentiable tensors, so the reference trainer's own loss code
(src/runner/trainers/acdc_vsr_refinenet_trainer.py:42-47,75-101: any torch loss on the output frames,
`loss.backward()`, any torch optimiser) works unchanged.  Forward and backward both run in libpvsr.so
(pvsr_plan_forward / pvsr_plan_backward); torch only carries the tensors.

Gradient semantics are the reference's: only the middle T frames carry gradients, everything produced at a warm-up
frame is a constant (refine_net.py:74-79,82-93,179-183), the inputs receive no gradient, and the never-applied
`refine_block.prelu.weight` gets grad None.
"""
import torch


class _RefineNetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, pl, names, *params):
        out = engine.run(pl).clone()          # the plan's output buffer is reused by the next step
        pl.fwd_serial = getattr(pl, 'fwd_serial', 0) + 1
        ctx.engine, ctx.pl, ctx.names, ctx.serial = engine, pl, names, pl.fwd_serial
        ctx.shape = out.shape
        # one differentiable tensor per output frame (views of `out`), list-major
        return tuple(out[l, t].unsqueeze(1) for l in range(out.shape[0]) for t in range(out.shape[1]))

    @staticmethod
    def backward(ctx, *grad_outs):
        engine, pl = ctx.engine, ctx.pl
        if pl.fwd_serial != ctx.serial:
            from .lib import PvsrError
            raise PvsrError('the activations saved by this forward were overwritten by a later forward of the same shape (the plan keeps ONE set of training buffers per shape): call backward before the next forward')
        n_lists, T = ctx.shape[0], ctx.shape[1]
        dout = pl.dout.view(n_lists * T, *ctx.shape[2:])
        for i, g in enumerate(grad_outs):
            if g is None:
                dout[i].zero_()
            else:
                dout[i].copy_(g.reshape(ctx.shape[2:]))
        bufs = engine.grad_buffers()
        for b in bufs.values():
            b.zero_()
        engine.backward(pl, bufs, generic=True)
        grads = []
        for k in ctx.names:
            # fresh tensors: autograd may keep them as .grad, the buffers are reused by the next backward
            grads.append(None if k == "refine_block.prelu.weight" else bufs[k].clone())
        return (None, None, None) + tuple(grads)


def refinenet_train_forward(net, inputs, pos_codes):
    """RefineNet.forward with autograd support (reference refine_net.py:61-135 under net.train())."""
    engine = net.engine
    pl = engine.train_plan(inputs)
    engine.stage_inputs(pl, [x.detach() for x in inputs], None if pos_codes is None else pos_codes.detach())
    named = list(net.named_parameters())
    names = tuple(k for k, _ in named)
    flat = _RefineNetFunction.apply(engine, pl, names, *[p for _, p in named])
    T = pl.T
    return tuple(list(flat[l * T:(l + 1) * T]) for l in range(pl.n_lists))
