"""python/shape_opt.py: the reconstruction loop (`optimize_shape`, :32-132) on top of the sdfrt pieces.

What the reference's loop does per iteration (:76-105) is kept one to one: per sensor `render` (primal at spp *
primal_spp_mult, adjoint re-render at spp with seed_grad = seed + 1 + n_sensors) -> loss / batch_size -> backward;
regulariser -> backward; validate_gradients; `opt.step()`; validate_params (x2 upsampling at the scheduled iterations,
LR schedule, max(sdf, bbox_sdf), re-distancing, EMA); push the parameters back into the scene.  Everything that touches
a grid runs on the GPU (render + adjoint, Adam step, regulariser, re-distancing, upsampling).  Scene loading from Mitsuba
XML, image files, videos and turntables are not part of this module: the caller passes a `Scene`, the variables and the
reference images (tensors)."""
import torch

from . import losses as losses_mod
from . import optimizers
from .render import render
from .scene import set_sensor_res

SDF_KEY = 'SamplingIntegrator.sdf.data'                      # python/constants.py:17
REFLECTANCE_KEY = 'main-bsdf.reflectance.volume.data'        # python/opt_configs.py:286


def update_scene(scene, opt):
    """`params.update(opt)` (python/shape_opt.py:55,105): bind the optimiser's tensors to the scene objects."""
    integrator = scene.integrator()
    if SDF_KEY in opt:
        integrator.sdf.data = opt[SDF_KEY]
        integrator.parameters_changed()
    if REFLECTANCE_KEY in opt:
        scene.albedo_grid = opt[REFLECTANCE_KEY]


def _ref_for(ref, width):
    """Reference image at the current rendering resolution: a tensor, or {width: tensor} (load_ref_images, :16-29)."""
    return ref[int(width)] if isinstance(ref, dict) else ref


def optimize_shape(scene, variables, ref_images, n_iter, spp=16, primal_spp_mult=4, learning_rate=4e-2, loss=None, batch_size=None,
                   param_averaging_beta=None, resolution_schedule=None, callback=None, opt=None):
    """Runs `n_iter` iterations and returns (opt, loss_values).

    variables          list of variables.Variable (at least the SdfVariable for SDF_KEY)
    ref_images         per sensor: (H, W, 3) CUDA tensor or {width: tensor}
    loss               loss(img, ref) -> scalar tensor; default losses.l1
    batch_size         sensors per iteration (default: all, in order); smaller batches are strided over the rig like
                       get_sensor_iterator (python/opt_configs.py:57-66)
    resolution_schedule {iteration: (resx, resy)}: the `update_scene` hook of the scene configs (python/opt_configs.py:53-55)
    callback           callback(i, loss_value, opt) after every iteration
    """
    sensors = scene.sensors()
    n_sensors = len(sensors)
    batch_size = n_sensors if batch_size is None else batch_size
    loss_fn = loss if loss is not None else losses_mod.l1
    sdf_object = scene.integrator().sdf
    if opt is None:
        opt = optimizers.Adam(lr=learning_rate)
        for v in variables:
            v.beta = param_averaging_beta if v.beta is None else v.beta
            v.initialize(opt)
    update_scene(scene, opt)
    seed = 0
    loss_values = []
    grad_clamp = {v.k: 1e-1 for v in variables if v.k == SDF_KEY}                       # SdfVariable.validate_gradient
    value_clamp = {v.k: v.value_clamp() for v in variables if getattr(v, 'value_clamp', None) and v.value_clamp() is not None}
    for i in range(n_iter):
        total = torch.zeros((), dtype=torch.float32, device=opt[SDF_KEY].device)
        opt.zero_grad()
        # python/opt_configs.py:57-66 (get_sensor_iterator): all sensors in order, or `batch_size` sensors strided over the rig
        # ((j * steps + i % steps) % n) so that every mini-batch covers the full range of view angles
        if batch_size >= n_sensors:
            batch = list(range(n_sensors))
        else:
            steps = -(-n_sensors // batch_size)
            batch = [(j * steps + i % steps) % n_sensors for j in range(batch_size)]
        for idx in batch:
            img = render(scene, sensor=idx, seed=seed, spp=spp * primal_spp_mult, seed_grad=seed + 1 + n_sensors, spp_grad=spp)
            seed += 1 + n_sensors
            view_loss = loss_fn(img, _ref_for(ref_images[idx], sensors[idx].film().crop_size()[0])) / batch_size
            view_loss.backward()
            total += view_loss.detach()
        for v in variables:
            reg = v.eval_regularizer(opt, sdf_object, i)
            if torch.is_tensor(reg) and reg.requires_grad:
                reg.backward()
                total += reg.detach()
        # validate_gradients + opt.step(): the gradient scrub/clamp and the value clamps are fused into the Adam kernel
        opt.step(grad_clamp=grad_clamp, value_clamp=value_clamp)
        for v in variables:
            v.validate(opt, i)
            v.update_mean(opt, i)
        if resolution_schedule and i in resolution_schedule:
            for s in sensors:
                set_sensor_res(s, resolution_schedule[i])
        update_scene(scene, opt)
        loss_values.append(float(total))
        if callback is not None:
            callback(i, loss_values[-1], opt)
    return opt, loss_values


def load_mean_parameters(scene, variables, opt):
    """python/opt_configs.py:143-146 + python/shape_opt.py:126-129: switch to the EMA of the parameters."""
    for v in variables:
        v.load_mean(opt)
        v.validate(opt, i=None)
    update_scene(scene, opt)
