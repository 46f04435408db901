"""`mi.render(scene, params, sensor, seed, spp, seed_grad, spp_grad)` contract:
primal at (seed, spp) without gradients; the backward pass re-renders at (seed_grad, spp_grad) and
accumulates into the `.grad` of the SDF (and reflectance) tensors."""
import torch


class _RenderOp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sdf_data, scene, sensor, seed, spp, seed_grad, spp_grad):
        integrator = scene.integrator()
        with torch.no_grad():
            image, film_a = integrator.render_with_adjoint_sweep(scene, sensor, seed, spp, seed_grad, spp_grad)
        ctx.args = (scene, film_a)
        return image

    @staticmethod
    def backward(ctx, grad_out):
        scene, film_a = ctx.args
        integrator = scene.integrator()
        data = integrator.sdf.data
        saved = data.grad
        data.grad = torch.zeros_like(data)
        # NOTE: the lane records of adjoint sweep 1 live in the context until the next sweep: backward() must follow the
        # forward of the same view before another view is rendered (the reference's loop does exactly that,
        # python/shape_opt.py:78-83)
        integrator.backprop_adjoint_sweep(scene, film_a, grad_out.contiguous())
        g, data.grad = data.grad, saved
        return g, None, None, None, None, None, None


def render(scene, params=None, sensor=0, seed=0, spp=0, seed_grad=0, spp_grad=0):
    integrator = scene.integrator()
    data = integrator.sdf.data
    if spp_grad == 0:
        spp_grad = spp
    if torch.is_grad_enabled() and data.requires_grad:
        return _RenderOp.apply(data, scene, sensor, seed, spp, seed_grad, spp_grad)
    return integrator.render(scene, sensor=sensor, seed=seed, spp=spp)
