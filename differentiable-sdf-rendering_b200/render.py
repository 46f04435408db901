"""`mi.render(scene, params, sensor, seed, spp, seed_grad, spp_grad)` contract:
primal at (seed, spp) without gradients; the backward pass re-renders at (seed_grad, spp_grad) and
accumulates into the `.grad` of the SDF (and reflectance) tensors."""
import torch


class _RenderOp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sdf_data, scene, sensor, seed, spp, seed_grad, spp_grad):
        integrator = scene.integrator()
        ctx.args = (scene, sensor, seed_grad, spp_grad)
        with torch.no_grad():
            return integrator.render(scene, sensor=sensor, seed=seed, spp=spp)

    @staticmethod
    def backward(ctx, grad_out):
        scene, sensor, seed_grad, spp_grad = ctx.args
        integrator = scene.integrator()
        data = integrator.sdf.data
        saved = data.grad
        data.grad = torch.zeros_like(data)
        integrator.render_backward(scene, None, grad_out.contiguous(), sensor=sensor, seed=seed_grad, spp=spp_grad)
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
