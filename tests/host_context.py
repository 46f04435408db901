"""TEST INFRASTRUCTURE: `HostContext` offers the interface of sdfrt.engine.Context on CPU tensors, computing with the kernels' per-lane
math compiled for the host (tests/hostmath.py) and the C / numpy oracles of the grid update.  It exists so that the Python mirror and the
compat layer (render op, view slots, variables, optimisation loop, the reference's own drivers) can be exercised by `-m "not gpu"` tests
on a machine without a GPU.  The product never uses it: sdfrt.engine.Context is the only context class the package ships, and it fails
without a CUDA device."""
import numpy as np
import torch

import hostmath as HM

INTEGRATOR_IDS = {'sdf_silhouette_reparam': 0, 'sdf_simple_shading_reparam': 1, 'sdf_direct_reparam': 2, 'sdf_prb_reparam': 3}


class _Sensor:
    width = 128
    height = 128


class HostContext:
    n_slots = 4

    def __init__(self, device=0):
        self.device = torch.device('cpu')
        self.sensor = _Sensor()
        self.sdf = None
        self.p = (0.0, 0.0, 0.0)
        self.to_world = None
        self.albedo = None
        self.tracer = {}
        self.warp = dict(kind=2, edge_eps=0.01, weight_strategy=6, clamping_thresh=0.05, normalize_warp_field=True, max_reparam_depth=-1,
                         n_aux_rays=16, kappa=1e5, power=3.0)
        self.sensor_args = None
        self.shading = dict(env_radiance=(1.0, 1.0, 1.0), env_dist=2.0 * 3.0 ** 0.5, albedo=(0.5, 0.5, 0.5), hide_emitters=False)
        self.integrator = 2
        self.antithetic = False
        self.use_mis = False
        self.envmap = None
        self.shadow_origin = 0
        self.prb = dict(max_depth=4, rr_depth=5)
        self.rank, self.world = 0, 1
        self.slot = 0
        self.views = [None] * self.n_slots       # (token, host scene, seed, spp, rec)
        self.generation = 0
        self.window = None
        self.n_forward = 0                        # adjoint sweeps run (tests count re-runs)

    # ---- scene state ---------------------------------------------------------------------------------------------------
    def set_sdf(self, data, p=(0.0, 0.0, 0.0)):
        self.sdf = data
        self.p = tuple(float(v) for v in p)
        self.sdf_shape = tuple(data.shape[:3])

    def set_sdf_transform(self, to_world=None):
        self.to_world = None if to_world is None else np.asarray(to_world, dtype=np.float64).reshape(4, 4)

    def set_albedo(self, grid):
        self.albedo = grid

    def set_tracer(self, **kw):
        self.tracer = dict(kw)

    def set_warp(self, kind=2, edge_eps=0.01, weight_strategy=6, clamping_thresh=0.05, normalize_warp_field=True, max_reparam_depth=-1,
                 n_aux_rays=16, kappa=1e5, power=3.0):
        self.warp = dict(kind=kind, edge_eps=edge_eps, weight_strategy=weight_strategy, clamping_thresh=clamping_thresh,
                         normalize_warp_field=normalize_warp_field, max_reparam_depth=max_reparam_depth, n_aux_rays=n_aux_rays,
                         kappa=kappa, power=power)

    def set_sensor(self, to_world, width, height, fov=39.0, near_clip=1e-2, far_clip=1e4, filter_radius=2.0, filter_stddev=0.5,
                   sample_border=True):
        self.sensor_args = (np.asarray(to_world, dtype=np.float64), int(width), int(height), fov, near_clip, far_clip, filter_radius,
                            filter_stddev, sample_border)
        self.sensor.width, self.sensor.height = int(width), int(height)

    def set_shading(self, env_radiance=(1.0, 1.0, 1.0), env_dist=2.0 * 3.0 ** 0.5, albedo=(0.5, 0.5, 0.5), hide_emitters=False):
        self.shading = dict(env_radiance=env_radiance, env_dist=env_dist, albedo=albedo, hide_emitters=hide_emitters)

    def set_integrator(self, name_or_id):
        self.integrator = INTEGRATOR_IDS[name_or_id] if isinstance(name_or_id, str) else int(name_or_id)

    def set_antithetic(self, enable):
        self.antithetic = bool(enable)

    def set_use_mis(self, enable):
        self.use_mis = bool(enable)

    def set_shadow_origin(self, mode):
        self.shadow_origin = int(mode)

    def set_envmap(self, rgb=None, scale=1.0, to_world=None):
        self.envmap = None if rgb is None else (np.array(rgb.detach().cpu().numpy() if torch.is_tensor(rgb) else rgb, dtype=np.float32), float(scale), to_world)

    def set_prb(self, max_depth=4, rr_depth=5):
        self.prb = dict(max_depth=int(max_depth), rr_depth=int(rr_depth))

    def set_partition(self, rank, world, rows=8):
        self.rank, self.world = rank, world
        HM.lib().hh_set_assign_rows(int(rows))

    def set_view_slot(self, slot):
        self.slot = int(slot)

    def num_view_slots(self):
        return self.n_slots

    def adjoint_token(self):
        v = self.views[self.slot]
        return 0 if v is None else v[0]

    def set_lane_window(self, lo=0, hi=0):
        self.window = None if (lo == 0 and hi == 0) else (int(lo), int(hi))

    def set_empty_space_skip(self, enable=True):
        pass

    # ---- the host scene of the current state -----------------------------------------------------------------------------
    def _scene(self):
        tw, W, H, fov, near, far, frad, fstd, border = self.sensor_args
        d = self.sdf.detach().cpu().numpy()
        d = d[..., 0] if d.ndim == 4 else d
        alb = None if self.albedo is None else self.albedo.detach().cpu().numpy()
        hs = HM.HostScene(d, tw, W, H, integrator=self.integrator, albedo_grid=alb, p=self.p, sdf_to_world=self.to_world)
        h = hs.h
        for k, v in self.tracer.items():
            setattr(h.tracer, k, v)
        h.warp.kind = self.warp['kind']
        h.warp.edge_eps = self.warp['edge_eps']
        h.warp.weight_strategy = self.warp['weight_strategy']
        h.warp.clamping_thresh = self.warp['clamping_thresh']
        h.warp.normalize_warp_field = int(self.warp['normalize_warp_field'])
        h.warp.max_reparam_depth = self.warp['max_reparam_depth']
        h.warp.n_aux_rays, h.warp.kappa, h.warp.power = self.warp.get('n_aux_rays', 16), self.warp.get('kappa', 1e5), self.warp.get('power', 3.0)
        h.sensor.fov_x_deg, h.sensor.near_clip, h.sensor.far_clip = fov, near, far
        h.sensor.filter_radius, h.sensor.filter_stddev, h.sensor.sample_border = frad, fstd, int(border)
        h.shading.env_radiance[:] = list(self.shading['env_radiance'])
        h.shading.env_dist = self.shading['env_dist']
        h.shading.albedo[:] = list(self.shading['albedo'])
        h.shading.hide_emitters = int(self.shading['hide_emitters'])
        h.antithetic = int(self.antithetic)
        h.use_mis = int(self.use_mis)
        if self.envmap is not None:
            if self.integrator == 3:
                raise RuntimeError("the envmap emitter is implemented for sdf_direct_reparam only")
            hs.set_envmap(*self.envmap)
        h.shadow_origin = self.shadow_origin
        h.prb.max_depth, h.prb.rr_depth = self.prb['max_depth'], self.prb['rr_depth']
        return hs

    # ---- rendering -------------------------------------------------------------------------------------------------------
    def film_dims(self):
        b = 2 if self.sensor_args[8] else 0
        return self.sensor.height + 2 * b, self.sensor.width + 2 * b, b

    def wavefront_size(self, spp):
        h, w, _ = self.film_dims()
        return h * w * spp

    def new_film(self):
        h, w, _ = self.film_dims()
        return torch.empty((h, w, 4), dtype=torch.float32)

    def render_primal(self, seed, spp, film=None):
        hs = self._scene()
        f = hs.prb_forward(seed, spp, self.rank, self.world)[0] if self.integrator == 3 else hs.forward(seed, spp, False, self.rank, self.world)[0]
        return torch.from_numpy(f)

    def develop(self, film, image=None):
        return torch.from_numpy(self._scene().develop(film.numpy()))

    def adjoint_accumulate(self, seed_grad, spp_grad, film=None):
        hs = self._scene()
        self.n_forward += 1
        if self.integrator == 3:
            f, rec = hs.prb_forward(seed_grad, spp_grad, self.rank, self.world)
        else:
            f, rec, _ = hs.forward(seed_grad, spp_grad, True, self.rank, self.world)
        self.generation += 1
        self.views[self.slot] = (self.generation, hs, seed_grad, spp_grad, rec, self.integrator == 3)
        return torch.from_numpy(f)

    def render_view(self, seed, spp, seed_grad, spp_grad, film_primal=None, film_adjoint=None):
        return self.render_primal(seed, spp), self.adjoint_accumulate(seed_grad, spp_grad)

    def adjoint_backprop(self, film, grad_image, grad_sdf, grad_albedo=None, grad_p=None):
        v = self.views[self.slot]
        if v is None:
            raise RuntimeError("sdfrt_adjoint_accumulate must run first")
        _, hs, seed, spp, rec, prb = v
        if prb:
            g, ga, gp, _ = hs.prb_backward(seed, spp, film.numpy(), rec, grad_image.detach().numpy(), self.rank, self.world)
        else:
            g, ga, gp, _ = hs.backward(seed, spp, film.numpy(), rec, grad_image.detach().numpy(), self.rank, self.world, lane_window=self.window)
        if grad_sdf is not None:
            grad_sdf += torch.from_numpy(g).reshape(grad_sdf.shape)
        if grad_albedo is not None and ga is not None:
            grad_albedo += torch.from_numpy(ga)
        if grad_p is not None:
            grad_p += torch.from_numpy(gp)

    def stats(self, reset=True):
        return {}

    # ---- lookups -----------------------------------------------------------------------------------------------------------
    def _grid_scene(self):
        d = self.sdf.detach().cpu().numpy()
        d = d[..., 0] if d.ndim == 4 else d
        return HM.HostScene(d, np.eye(4), 8, 8, p=self.p, sdf_to_world=self.to_world)

    def eval_sdf(self, x, order=0):
        return torch.from_numpy(self._grid_scene().eval_sdf(x.detach().cpu().numpy(), order))

    def trace_rays(self, o, d, maxt=None, differentiable=True):
        hs = self._grid_scene()
        for k, v in self.tracer.items():
            setattr(hs.h.tracer, k, v)
        out = hs.trace(o.numpy(), d.numpy(), None if maxt is None else maxt.numpy(), differentiable)
        return {k: torch.from_numpy(v) for k, v in out.items()}

    # ---- grid update: the CPU oracles ----------------------------------------------------------------------------------------
    def redistance(self, phi, lower_bound=None, return_iters=False):
        from oracle import redistance_oracle as ro
        shape = phi.shape
        g = phi.detach().cpu().numpy().reshape(shape[:3]).astype(np.float32)
        if lower_bound is not None:
            g = np.maximum(g, lower_bound.detach().cpu().numpy().reshape(g.shape).astype(np.float32))
        out = torch.from_numpy(np.ascontiguousarray(ro.redistance(g), dtype=np.float32)).reshape(shape)
        return (out, 0) if return_iters else out

    def redistance_status(self):
        pass

    def adam_step(self, param, grad, m, v, lr, t, beta1=0.9, beta2=0.999, eps=1e-8, grad_clamp=0.0, lo=1.0, hi=0.0):
        from oracle import grid_update_oracle as GO
        p2, g2, m2, v2 = GO.adam_step(param.detach().numpy(), grad.detach().numpy(), m.numpy(), v.numpy(), lr, t, beta1, beta2, eps, grad_clamp, lo, hi)
        with torch.no_grad():
            param.copy_(torch.from_numpy(p2.astype(np.float32)))
            grad.copy_(torch.from_numpy(g2.astype(np.float32)))
            m.copy_(torch.from_numpy(m2.astype(np.float32)))
            v.copy_(torch.from_numpy(v2.astype(np.float32)))

    def laplacian_reg(self, data, weight=1.0, grad=None, return_loss=True):
        from oracle import grid_update_oracle as GO
        val, g = GO.laplacian_reg(data.detach().numpy())
        if grad is not None:
            grad += weight * torch.from_numpy(g.astype(np.float32)).reshape(grad.shape)
        return torch.tensor(weight * val, dtype=torch.float64) if return_loss else None


def install(monkeypatch):
    """Route the Python mirror to a HostContext (pytest monkeypatch: undone after the test)."""
    import sdfrt
    shared = {}

    def shared_context(device=0):
        if 0 not in shared:
            shared[0] = HostContext(0)
        return shared[0]
    monkeypatch.setattr(sdfrt.engine, 'Context', HostContext)
    monkeypatch.setattr(sdfrt.redistancing, 'shared_context', shared_context)
    monkeypatch.setattr(sdfrt, 'Context', HostContext)
    return shared_context
