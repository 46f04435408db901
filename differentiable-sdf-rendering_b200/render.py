"""`mi.render(scene, params, sensor, seed, spp, seed_grad, spp_grad)` contract:
primal at (seed, spp) without gradients; the backward pass re-renders at (seed_grad, spp_grad) and
back-propagates into the SDF grid and the reflectance grid (both are differentiable inputs of the op)."""
import torch


class _RenderOp(torch.autograd.Function):
    """Mitsuba's `_RenderOp` (python/shape_opt.py:78-83 calls it through mi.render): `eval` = primal render, `backward` =
    integrator.render_backward.  Here the re-render of render_backward (adjoint sweep 1) is hoisted next to the primal render (it
    does not depend on the loss) and its per-lane tracer records stay inside the sdfrt context until `backward`.  The context
    has a few view slots (sdfrt_set_view_slot), used round robin: the op remembers its slot and that sweep's token and, if the slot was
    reused in the meantime (a loss summed over more views than there are slots), re-runs sweep 1 of its own view before
    back-propagating -- the result is the same either way."""

    @staticmethod
    def forward(ctx, sdf_data, albedo, scene, sensor, seed, spp, seed_grad, spp_grad):
        integrator = scene.integrator()
        with torch.no_grad():
            image, film_a = integrator.render_with_adjoint_sweep(scene, sensor, seed, spp, seed_grad, spp_grad)
        ctx.view = (scene, sensor, seed_grad, spp_grad, film_a, integrator._view_slot, integrator.sdf.ctx.adjoint_token())
        ctx.shapes = (sdf_data.shape, None if albedo is None else albedo.shape)
        return image

    @staticmethod
    def backward(ctx, grad_out):
        scene, sensor, seed_grad, spp_grad, film_a, slot, token = ctx.view
        integrator = scene.integrator()
        c = integrator.sdf.ctx
        c.set_view_slot(slot)
        if token == 0 or c.adjoint_token() != token:
            film_a = integrator.adjoint_sweep(scene, sensor, seed_grad, spp_grad)
        dev = grad_out.device
        g_sdf = torch.zeros(ctx.shapes[0], dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
        g_alb = torch.zeros(ctx.shapes[1], dtype=torch.float32, device=dev) if (ctx.shapes[1] is not None and ctx.needs_input_grad[1]) else None
        integrator.backprop_into(film_a, grad_out.contiguous(), g_sdf, g_alb)
        return g_sdf, g_alb, None, None, None, None, None, None


def render(scene, params=None, sensor=0, seed=0, spp=0, seed_grad=0, spp_grad=0):
    integrator = scene.integrator()
    data = integrator.sdf.data
    alb = scene.albedo_grid
    if spp_grad == 0:
        spp_grad = spp
    wants_grad = data.requires_grad or (alb is not None and alb.requires_grad)
    if torch.is_grad_enabled() and wants_grad:
        return _RenderOp.apply(data, alb, scene, sensor, seed, spp, seed_grad, spp_grad)
    return integrator.render(scene, sensor=sensor, seed=seed, spp=spp)
