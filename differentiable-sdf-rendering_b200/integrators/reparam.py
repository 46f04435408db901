"""Host-side mirror of python/integrators/reparam.py: same attributes and method names; `render`
and `render_backward` launch the CUDA passes through the C ABI."""
import torch

from .. import engine
from ..shapes import Grid3d
from ..warp import warp_params_of


class Properties(dict):
    """Stand-in for mi.Properties: a dict with .get / .has_property."""

    def has_property(self, k):
        return k in self


class ReparamIntegrator:
    integrator_name = None

    def __init__(self, props=None, sdf=None):
        props = Properties(props or {})
        self.max_depth = props.get('max_depth', 4)
        self.weight_by_spp = props.get('weight_by_spp', False)
        assert not self.weight_by_spp, "Not supported"  # python/integrators/reparam.py:16
        self.use_mis = props.get('use_mis', False)   # python/integrators/reparam.py:17; used by sdf_direct_reparam (:76-79,87-107)
        self.antithetic_sampling = props.get('antithetic_sampling', False)   # python/integrators/reparam.py:19,167-178
        self.use_aovs = props.get('use_aovs', False)
        self.hide_emitters = props.get('hide_emitters', False)
        # python/integrators/reparam.py:21-30: `sdf_filename` (+ optional `sdf_to_world`) build the Grid3d; a ready Grid3d can
        # be passed instead (what shape_opt does after `scene.integrator().sdf`)
        if sdf is None and props.get('sdf_filename', '') != '':
            tf = props.get('sdf_to_world', None) if props.has_property('sdf_to_world') else None
            sdf = Grid3d(props['sdf_filename'], transform=tf)
        self.sdf = sdf
        self.warp_field = None
        self.last_film = None

    # ---- parameter plumbing (python/integrators/reparam.py:269-277) ----
    def traverse(self, cb):
        if self.sdf is not None:
            self.sdf.traverse(cb)

    def parameters_changed(self, keys=None):
        if self.sdf is not None:
            self.sdf.parameters_changed(keys)

    # ---- rendering ----
    def _bind(self, scene, sensor):
        if self.sdf is None:
            raise ValueError("integrator has no SDF")
        if isinstance(sensor, int):
            sensor = scene.sensors()[sensor]
        ctx = self.sdf.ctx
        # the grid tensor may have been updated in place since the last view (optimiser step): rebind it so that the library rebuilds
        # its derived copies -- the `parameters_changed` -> `set_tensor` of python/shapes.py:477-479
        ctx.set_sdf(self.sdf.data, self.sdf.p)
        film = sensor.film()
        ctx.set_sensor(sensor.to_world, film.width, film.height, sensor.fov, sensor.near_clip, sensor.far_clip,
                       film.filter_radius, film.filter_stddev, film.sample_border)
        ctx.set_shading(scene.env_radiance, scene.env_dist, scene.albedo, self.hide_emitters)
        ctx.set_albedo(scene.albedo_grid)
        env_key = (id(scene), getattr(scene, 'envmap_version', 0)) if getattr(scene, 'envmap', None) is not None else None
        if getattr(ctx, '_envmap_key', None) != env_key:   # the bitmap is uploaded once per scene / change, not per view
            ctx.set_envmap(None if env_key is None else scene.envmap, getattr(scene, 'envmap_scale', 1.0), getattr(scene, 'envmap_to_world', None))
            ctx._envmap_key = env_key
        ctx.set_integrator(self.integrator_name)
        ctx.set_antithetic(self.antithetic_sampling)
        ctx.set_use_mis(self.use_mis and self.integrator_name == 'sdf_direct_reparam')
        ctx.set_warp(**warp_params_of(self.warp_field))
        ctx.set_tracer(**self.sdf.tracer_params())
        return ctx

    def _check_wavefront(self, ctx, spp):
        n = ctx.wavefront_size(spp)
        if n > 0xffffffff:  # python/integrators/reparam.py:48-50
            raise Exception(f"Wavefront {n} exceeds {0xffffffff}")

    def render(self, scene, sensor=0, seed=0, spp=0, develop=True, evaluate=True):
        if not develop:
            raise Exception("Must use develop=True for this AD integrator")  # python/integrators/reparam.py:122-123
        ctx = self._bind(scene, sensor)
        self._check_wavefront(ctx, spp)
        film = ctx.render_primal(seed, spp)
        return ctx.develop(film)

    def render_with_adjoint_sweep(self, scene, sensor, seed, spp, seed_grad, spp_grad):
        """Primal image plus the film of adjoint sweep 1, from ONE pipelined call (sdfrt_render_view): what `mi.render`
        does when gradients are enabled, with the re-render of render_backward hoisted next to the primal render (it does
        not depend on the loss).  `backprop_adjoint_sweep` then plays the role of render_backward."""
        ctx = self.sdf.ctx
        # adjoint state goes to the next view slot (round robin): up to num_view_slots() views can be rendered before their backward
        self._view_slot = (getattr(self, '_view_slot', -1) + 1) % ctx.num_view_slots()
        ctx.set_view_slot(self._view_slot)
        ctx = self._bind(scene, sensor)
        self._check_wavefront(ctx, spp)
        self._check_wavefront(ctx, spp_grad)
        film, film_a = ctx.render_view(seed, spp, seed_grad, spp_grad)
        self.last_film = film_a
        return ctx.develop(film), film_a

    def adjoint_sweep(self, scene, sensor, seed_grad, spp_grad):
        """Adjoint sweep 1 alone (the re-render of render_backward): rebinds the view and refills the context's lane records."""
        ctx = self._bind(scene, sensor)   # (into the currently selected view slot)
        self._check_wavefront(ctx, spp_grad)
        self.last_film = ctx.adjoint_accumulate(seed_grad, spp_grad)
        return self.last_film

    def backprop_into(self, film_a, grad_in, grad_sdf, grad_albedo):
        """Adjoint sweep 2 for the view whose records the context holds: ADDS into grad_sdf / grad_albedo (either may be None)
        and into `sdf.p_grad` (the gradient w.r.t. the translation `sdf.p`, python/shapes.py:389)."""
        ctx = self.sdf.ctx
        if not hasattr(self.sdf, 'p_grad') or self.sdf.p_grad is None:
            self.sdf.p_grad = torch.zeros(3, dtype=torch.float32, device=ctx.device)
        ctx.adjoint_backprop(film_a, grad_in.contiguous(), grad_sdf, grad_albedo, self.sdf.p_grad)

    def backprop_adjoint_sweep(self, scene, film_a, grad_in):
        data = self.sdf.data
        if data.grad is None:
            data.grad = torch.zeros_like(data)
        galb = None
        if scene.albedo_grid is not None and scene.albedo_grid.requires_grad:
            if scene.albedo_grid.grad is None:
                scene.albedo_grid.grad = torch.zeros_like(scene.albedo_grid)
            galb = scene.albedo_grid.grad
        self.backprop_into(film_a, grad_in, data.grad, galb)

    def render_backward(self, scene, params, grad_in, sensor=0, seed=0, spp=0):
        """Accumulates dLoss/d(sdf.data) into `sdf.data.grad` (and the reflectance grid's .grad)."""
        ctx = self._bind(scene, sensor)
        self._check_wavefront(ctx, spp)
        film = ctx.adjoint_accumulate(seed, spp)
        data = self.sdf.data
        if data.grad is None:
            data.grad = torch.zeros_like(data)
        galb = None
        if scene.albedo_grid is not None and scene.albedo_grid.requires_grad:
            if scene.albedo_grid.grad is None:
                scene.albedo_grid.grad = torch.zeros_like(scene.albedo_grid)
            galb = scene.albedo_grid.grad
        if not hasattr(self.sdf, 'p_grad') or self.sdf.p_grad is None:
            self.sdf.p_grad = torch.zeros(3, dtype=torch.float32, device=ctx.device)
        ctx.adjoint_backprop(film, grad_in.contiguous(), data.grad, galb, self.sdf.p_grad)
        self.last_film = film
