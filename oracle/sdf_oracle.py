"""CPU oracle for the differentiable-SDF-rendering hot path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference (rgl-epfl/differentiable-sdf-rendering) ships no tests,
golden vectors or scenes, and its runtime (mitsuba 3.0-3.5 / drjit <= 0.4, unpinned,
`README.md:48-50`) is not installable offline.  This file is therefore a *restatement*:
every function follows the cited reference lines, with Dr.Jit's AD replaced by
torch.autograd on the CPU (`replace_grad(a, b)` == `a.detach() + (b - b.detach())`,
`dr.detach` == `.detach()`, `dr.suspend_grad` == `torch.no_grad`).  The third-party
pieces (cubic texture, PCG32/TEA sampler, perspective sensor, gaussian film, constant
emitter, diffuse BSDF, spawn_ray_to) are restated from the published Mitsuba 3 /
Dr.Jit algorithms as listed in SURVEY.md App. B and oracle/ref_semantics.md.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path (libsdfrt.so) never does.

All ray data are torch tensors of shape (N,) or (N,3); dtype is selectable (float64 is
the gold used for parity, float32 mirrors the arithmetic type of the CUDA path).
"""
import math
from dataclasses import dataclass, field

import numpy as np
import torch

INF = float('inf')

# --------------------------------------------------------------------------------------
# small AD helpers (Dr.Jit semantics)
# --------------------------------------------------------------------------------------


def replace_grad(a, b):
    """dr.replace_grad: value of `a`, derivative of `b`."""
    a = a.detach() if torch.is_tensor(a) else torch.as_tensor(a, dtype=b.dtype)
    return a + (b - b.detach())


def dr_sign(x):
    """dr.sign: +1 for x >= 0, -1 otherwise."""
    return torch.where(x >= 0, torch.ones_like(x), -torch.ones_like(x))


def dot(a, b):
    return (a * b).sum(-1)


def normalize(v):
    return v / torch.linalg.norm(v, dim=-1, keepdim=True)


def outer(v, w):
    """math_util.outer_product (python/math_util.py:21-25)."""
    return v[..., :, None] * w[..., None, :]


def eye3(like):
    return torch.eye(3, dtype=like.dtype).expand(like.shape[:-1] + (3, 3))


def normalize_sqr(x):
    """math_util.normalize_sqr (python/math_util.py:13-18): x/|x|^2 and its Jacobian."""
    x2 = dot(x, x)
    jac = eye3(x) / x2[..., None, None] - (2 / (x2 * x2))[..., None, None] * outer(x, x)
    return x / x2[..., None], jac


def bbox_distance_inside_d(x, bmin, bmax):
    """math_util.bbox_distance_inside_d (python/math_util.py:31-41)."""
    d = torch.minimum((x - bmin).min(-1).values, (bmax - x).min(-1).values).clamp_min(0.0)
    bbox_max_dist = (bmax - x).abs()
    bbox_min_dist = (bmin - x).abs()
    md = torch.minimum(bbox_min_dist, bbox_max_dist)
    n = torch.zeros_like(x)
    mx, my, mz = md[..., 0], md[..., 1], md[..., 2]
    n[(mx < my) & (mx < mz)] = torch.tensor([1.0, 0.0, 0.0], dtype=x.dtype)
    n[(my < mz) & (my < mx)] = torch.tensor([0.0, 1.0, 0.0], dtype=x.dtype)
    n[(mz < mx) & (mz < my)] = torch.tensor([0.0, 0.0, 1.0], dtype=x.dtype)
    d_d = torch.where((d > 0)[..., None], n * dr_sign(bbox_max_dist - bbox_min_dist), torch.zeros_like(x))
    return d, d_d


# --------------------------------------------------------------------------------------
# Sampler: Mitsuba `independent` = PCG32 seeded through TEA
# --------------------------------------------------------------------------------------

_M32 = np.uint64(0xFFFFFFFF)
PCG32_MULT = np.uint64(0x5851F42D4C957F2D)


def sample_tea_32(v0, v1, rounds=4):
    """Tiny Encryption Algorithm, 4 rounds, as mitsuba `sample_tea_32`."""
    v0 = np.asarray(v0, dtype=np.uint64) & _M32
    v1 = np.asarray(v1, dtype=np.uint64) & _M32
    s = np.uint64(0)
    for _ in range(rounds):
        s = (s + np.uint64(0x9E3779B9)) & _M32
        v0 = (v0 + ((((v1 << np.uint64(4)) & _M32) + np.uint64(0xA341316C)) ^ (v1 + s) ^ ((v1 >> np.uint64(5)) + np.uint64(0xC8013EA4)))) & _M32
        v1 = (v1 + ((((v0 << np.uint64(4)) & _M32) + np.uint64(0xAD90777D)) ^ (v0 + s) ^ ((v0 >> np.uint64(5)) + np.uint64(0x7E95761E)))) & _M32
    return v0, v1


class PCG32:
    """drjit PCG32 (XSH-RR 64/32) vectorised over lanes."""

    def __init__(self, initstate, initseq):
        with np.errstate(over='ignore'):
            self.state = np.zeros_like(initstate, dtype=np.uint64)
            self.inc = ((initseq.astype(np.uint64) << np.uint64(1)) | np.uint64(1))
            self.next_u32()
            self.state = self.state + initstate.astype(np.uint64)
            self.next_u32()

    def next_u32(self, mask=None):
        """mask: lanes that advance (drjit `next_uint32(mask)`); the others keep their state and return garbage."""
        with np.errstate(over='ignore'):
            old = self.state
            new = old * PCG32_MULT + self.inc
            self.state = new if mask is None else np.where(mask, new, old)
            xorshifted = (((old >> np.uint64(18)) ^ old) >> np.uint64(27)) & _M32
            rot = (old >> np.uint64(59)) & _M32
            res = ((xorshifted >> rot) | (xorshifted << ((np.uint64(32) - rot) & np.uint64(31)))) & _M32
        return res.astype(np.uint32)

    def next_f32(self, mask=None):
        u = self.next_u32(mask)
        bits = (u >> np.uint32(9)) | np.uint32(0x3F800000)
        return bits.view(np.float32) - np.float32(1.0)


class IndependentSampler:
    """`sampler.seed(seed, wavefront)` + next_1d/next_2d (python/integrators/reparam.py:39-51)."""

    def __init__(self, seed, wavefront, dtype, lane_offset=0, lane_count=None):
        idx = np.arange(lane_offset, lane_offset + (wavefront if lane_count is None else lane_count), dtype=np.uint64)
        v0, v1 = sample_tea_32(np.full_like(idx, np.uint64(seed & 0xFFFFFFFF)), idx)
        self.rng = PCG32(v0, v1)
        self.dtype = dtype

    def next_1d(self, active=None):
        """active: (N,) bool tensor / array; Mitsuba's samplers only advance the active lanes (`next_1d(active)`)."""
        m = None if active is None else np.asarray(active, dtype=bool)
        return torch.from_numpy(self.rng.next_f32(m).copy()).to(self.dtype)

    def next_2d(self, active=None):
        a = self.next_1d(active)
        b = self.next_1d(active)
        return torch.stack([a, b], -1)


# --------------------------------------------------------------------------------------
# Tricubic B-spline texture (drjit Texture3f.eval_cubic / _grad / _hessian), SURVEY App. B
# --------------------------------------------------------------------------------------


def _bspline_w(a):
    a2 = a * a
    a3 = a2 * a
    w = torch.stack([(-a3 + 3 * a2 - 3 * a + 1) / 6, (3 * a3 - 6 * a2 + 4) / 6,
                     (-3 * a3 + 3 * a2 + 3 * a + 1) / 6, a3 / 6], -1)
    dw = torch.stack([-0.5 * a2 + a - 0.5, 1.5 * a2 - 2 * a, -1.5 * a2 + a + 0.5, 0.5 * a2], -1)
    ddw = torch.stack([1 - a, 3 * a - 2, -3 * a + 1, a], -1)
    return w, dw, ddw


class Grid3d:
    """python/shapes.py:375-483 + SDFBase knobs `:28-41`.  `to_world`: optional affine 4x4 (Grid3d's `transform`)."""

    def __init__(self, data, p=None, to_world=None):
        data = torch.as_tensor(data)
        if data.dim() == 4:
            data = data[..., 0]
        self.data = data  # (Z, Y, X), x fastest; may require grad
        self.dtype = data.dtype
        self.p = torch.zeros(3, dtype=self.dtype) if p is None else torch.as_tensor(p, dtype=self.dtype)
        self.has_transform = to_world is not None                      # python/shapes.py:381
        self.to_world = np.eye(4) if to_world is None else np.asarray(to_world, dtype=np.float64).reshape(4, 4)
        self.to_local = np.linalg.inv(self.to_world)
        # SDFBase.__init__ defaults (python/shapes.py:28-41)
        self.step_scale = 1.0
        self.trace_eps = 1e-6
        self.refine_intersection = True
        self.use_extra_weight = True
        self.extra_thresh = 0.05
        self.sil_weight_offset = 0.05
        self.sil_weight_epsilon = 1e-6
        self.weight_power = 3
        self.use_weight_sum_weight = True
        self.max_steps = 1 << 16  # safety cap (reference has none, python/shapes.py:170)
        self.n_evals = 0          # 64-tap evaluation sites (for the roofline byte count)

    @property
    def shape(self):
        return tuple(self.data.shape)

    def bbox(self):
        """python/shapes.py:394-403,416-418: AABB of the transformed unit cube +-0.05; ignores self.p (Appendix C #13)."""
        c = np.array([[x, y, z, 1.0] for x in (0, 1) for y in (0, 1) for z in (0, 1)]) @ self.to_world.T
        lo = torch.as_tensor(c[:, :3].min(0) - 0.05, dtype=self.dtype)
        hi = torch.as_tensor(c[:, :3].max(0) + 0.05, dtype=self.dtype)
        return lo, hi

    def _taps(self, x, order):
        """Direct 64-tap evaluation; order 0: value, 1: +gradient, 2: +Hessian."""
        self.n_evals += x.shape[0]
        Z, Y, X = self.data.shape
        res = torch.tensor([X, Y, Z], dtype=self.dtype)
        xl = x - self.p  # call_wrap: to_local @ (x - p) (python/shapes.py:408-414)
        A = None
        if self.has_transform:
            A = torch.as_tensor(self.to_local[:3, :3], dtype=self.dtype)
            xl = xl @ A.T + torch.as_tensor(self.to_local[:3, 3], dtype=self.dtype)
        u = xl * res - 0.5
        uf = torch.floor(u.detach())
        a = u - uf
        i0 = uf.to(torch.int64) - 1
        wx, dwx, ddwx = _bspline_w(a[:, 0])
        wy, dwy, ddwy = _bspline_w(a[:, 1])
        wz, dwz, ddwz = _bspline_w(a[:, 2])
        o = torch.arange(4)
        ix = (i0[:, 0:1] + o).clamp(0, X - 1)
        iy = (i0[:, 1:2] + o).clamp(0, Y - 1)
        iz = (i0[:, 2:3] + o).clamp(0, Z - 1)
        lin = (iz[:, :, None, None] * Y + iy[:, None, :, None]) * X + ix[:, None, None, :]
        f = self.data.reshape(-1)[lin.reshape(-1)].reshape(-1, 4, 4, 4)  # [n, z, y, x]

        def red(az, ay, ax):
            return torch.einsum('nzyx,nz,ny,nx->n', f, az, ay, ax)

        v = red(wz, wy, wx)
        if order == 0:
            return v
        g = torch.stack([red(wz, wy, dwx) * res[0], red(wz, dwy, wx) * res[1], red(dwz, wy, wx) * res[2]], -1)
        if A is not None:
            g = g @ A  # to_world @ Normal3f(g) = to_local^T g (python/shapes.py:426-428,434-436,445-447)
        if order == 1:
            return v, g
        hxx = red(wz, wy, ddwx) * res[0] * res[0]
        hyy = red(wz, ddwy, wx) * res[1] * res[1]
        hzz = red(ddwz, wy, wx) * res[2] * res[2]
        hxy = red(wz, dwy, dwx) * res[0] * res[1]
        hxz = red(dwz, wy, dwx) * res[0] * res[2]
        hyz = red(dwz, dwy, wx) * res[1] * res[2]
        H = torch.stack([torch.stack([hxx, hxy, hxz], -1), torch.stack([hxy, hyy, hyz], -1),
                         torch.stack([hxz, hyz, hzz], -1)], -2)
        if A is not None:
            H = A.T @ H @ A  # python/shapes.py:448
        return v, g, H

    def eval(self, x):
        """python/shapes.py:420-421 (value only)."""
        return self._taps(x, 0)

    def eval_and_grad(self, x):
        """python/shapes.py:430-436."""
        return self._taps(x, 1)

    def eval_grad(self, x):
        """python/shapes.py:423-428."""
        return self._taps(x, 1)[1]

    def eval_all(self, x):
        """python/shapes.py:438-450: (v, detach v, g, detach g, H)."""
        v, g, H = self._taps(x, 2)
        return v, v.detach(), g, g.detach(), H

    # ---- python/shapes.py:68-113 -----------------------------------------------------
    def eval_trace_weight(self, d, i, bmin, bmax, x, sdf_value, sdf_grad, hessian):
        n_dot_d = dot(sdf_grad, d)
        n_dot_n = dot(sdf_grad, sdf_grad)
        dot_ratio = n_dot_d / n_dot_n
        denom = self.sil_weight_epsilon + sdf_value.abs() + self.sil_weight_offset * n_dot_d * dot_ratio
        dist_weight = 1 / denom ** self.weight_power
        bbox_dist, bbox_dist_d = bbox_distance_inside_d(x, bmin, bmax)
        bbox_eps = 0.01
        bbox_weight = torch.where(i > 0, torch.minimum(bbox_dist, torch.full_like(bbox_dist, bbox_eps)) / bbox_eps,
                                  torch.ones_like(bbox_dist))
        weight = dist_weight * bbox_weight
        bbox_weight_d = torch.where(((i > 0) & (bbox_dist < bbox_eps))[:, None], bbox_dist_d / bbox_eps,
                                    torch.zeros_like(bbox_dist_d))
        gradient = 2 * dot_ratio[:, None] * (d - dot_ratio[:, None] * sdf_grad)
        denom_d = dr_sign(sdf_value)[:, None] * sdf_grad + self.sil_weight_offset * torch.einsum('ni,nij->nj', gradient, hessian)
        dist_weight_d = (-self.weight_power * dist_weight / denom)[:, None] * denom_d
        weight_d = dist_weight[:, None] * bbox_weight_d + bbox_weight[:, None] * dist_weight_d
        return weight, weight_d

    # ---- AABB slab test (mitsuba BoundingBox3f::ray_intersect), SURVEY App. B --------
    @staticmethod
    def _aabb(o, d, bmin, bmax):
        d_rcp = 1.0 / d
        t1 = (bmin - o) * d_rcp
        t2 = (bmax - o) * d_rcp
        t1p = torch.minimum(t1, t2)
        t2p = torch.maximum(t1, t2)
        mint = t1p.max(-1).values
        maxt = t2p.min(-1).values
        active = ((d != 0) | ((o >= bmin) & (o <= bmax))).all(-1) & (maxt >= mint)
        inside = ((o >= bmin) & (o <= bmax)).all(-1)
        return active, mint, maxt, inside

    # ---- python/shapes.py:115-288: the differentiable tracer --------------------------
    @torch.no_grad()
    def ray_intersect(self, o, d, ray_maxt, active, extra_outputs=None):
        N = o.shape[0]
        dt = self.dtype
        d = normalize(d)
        bmin, bmax = self.bbox()
        intersects_bbox, mint, maxt, inside_bbox = self._aabb(o, d, bmin, bmax)
        intersects_bbox = intersects_bbox & ((mint > 0) | inside_bbox)
        active = active & intersects_bbox
        maxt = torch.minimum(maxt, ray_maxt)
        trace_eps = self.trace_eps * maxt.clamp_min(1.0)
        its_t = torch.full((N,), INF, dtype=dt)
        t = torch.where(inside_bbox, torch.zeros_like(mint), mint + 1e-5)
        warp_t = torch.zeros(N, dtype=dt)
        prev_surf_dist = torch.zeros(N, dtype=dt)
        prev_sdf_grad_c = torch.zeros(N, 3, dtype=dt)
        weight_sum = torch.zeros(N, dtype=dt)
        mixed_sum_d = torch.zeros(N, 3, dtype=dt)
        weight_d_sum = torch.zeros(N, 3, dtype=dt)
        it = torch.zeros(N, dtype=torch.int64)
        extra_weight_sum = torch.zeros(N, dtype=dt)
        extra_weight_sum_d = torch.zeros(N, 3, dtype=dt)

        # entry face normal + initial t_d (python/shapes.py:155-164)
        bbox_its_p = o + t[:, None] * d
        md = torch.minimum((bmin - bbox_its_p).abs(), (bmax - bbox_its_p).abs())
        n = torch.zeros(N, 3, dtype=dt)
        n[(md[:, 0] < md[:, 1]) & (md[:, 0] < md[:, 2])] = torch.tensor([1.0, 0.0, 0.0], dtype=dt)
        n[(md[:, 1] < md[:, 2]) & (md[:, 1] < md[:, 0])] = torch.tensor([0.0, 1.0, 0.0], dtype=dt)
        n[(md[:, 2] < md[:, 0]) & (md[:, 2] < md[:, 1])] = torch.tensor([0.0, 0.0, 1.0], dtype=dt)
        d_dot_n = dot(d, n)
        t_d = torch.zeros(N, 3, dtype=dt)
        m0 = (~inside_bbox) & (d_dot_n.abs() > 0)
        t_d[m0] = (-n[m0] / d_dot_n[m0, None]) * t[m0, None]

        act = torch.nonzero(active).squeeze(-1)
        step = 0
        while act.numel() > 0 and step < self.max_steps:
            step += 1
            a = act
            da = d[a]
            ta = t[a]
            x = o[a] + ta[:, None] * da
            sdf_value, sdf_grad, hessian = self._taps(x, 2)
            if self.step_scale != 1.0:
                sdf_value = sdf_value * self.step_scale
                sdf_grad = sdf_grad * self.step_scale
                hessian = hessian * self.step_scale
            intersected = sdf_value < trace_eps[a]
            its_t[a] = torch.where(intersected, ta, its_t[a])
            surf_dist = sdf_value.abs()
            weight, weight_d = self.eval_trace_weight(da, it[a], bmin, bmax, x, sdf_value, sdf_grad, hessian)

            ews = extra_weight_sum[a]
            psd = prev_surf_dist[a]
            if self.use_extra_weight:
                inv_extra_w_den = 1 / torch.minimum(torch.full_like(surf_dist, self.extra_thresh), surf_dist)
                dist_difference = psd - surf_dist
                ews = ews + torch.where(dist_difference >= 0, dist_difference * inv_extra_w_den, torch.zeros_like(ews))
                ews = torch.minimum(ews, torch.ones_like(ews))
            curr_segment_value = torch.where(intersected, torch.zeros_like(surf_dist), surf_dist)
            segment_length = 0.5 * (curr_segment_value + psd)
            weight_increment = segment_length * weight
            if self.use_extra_weight:
                weight_increment = weight_increment * ews
            weight_sum[a] = weight_sum[a] + weight_increment
            warp_t[a] = warp_t[a] + weight_increment * ta

            tda = t_d[a]

            def convert_deriv(in_d):
                return ta[:, None] * in_d + dot(da, in_d)[:, None] * tda

            weight_d = convert_deriv(weight_d)
            sdf_grad_c = convert_deriv(sdf_grad)
            pg = prev_sdf_grad_c[a]
            segment_d = 0.5 * (sdf_grad_c + pg)
            if self.use_extra_weight:
                sdf_sign = dr_sign(sdf_value)
                surf_dist_d = sdf_sign[:, None] * sdf_grad_c
                extra_w_d = (pg - surf_dist_d) * inv_extra_w_den[:, None]
                extra_w_d = extra_w_d - (dist_difference * inv_extra_w_den ** 2)[:, None] * torch.where(
                    (sdf_value < self.extra_thresh)[:, None], surf_dist_d, torch.zeros_like(surf_dist_d))
                ewd = extra_weight_sum_d[a] + torch.where((dist_difference > 0.0)[:, None], extra_w_d, torch.zeros_like(extra_w_d))
                ewd = torch.where(((ews >= 1.0) | (ews <= 0.0))[:, None], torch.zeros_like(ewd), ewd)
                extra_weight_sum_d[a] = ewd
                extra_weight_sum[a] = ews
                weight_d = weight[:, None] * ewd + weight_d * ews[:, None]
                weight = weight * ews
            weight_increment_d = weight[:, None] * segment_d + weight_d * segment_length[:, None]
            mixed_sum_d[a] = mixed_sum_d[a] + weight_increment_d * ta[:, None] + (weight * segment_length)[:, None] * tda
            t_d[a] = tda + sdf_grad_c
            weight_d_sum[a] = weight_d_sum[a] + weight_increment_d
            it[a] = it[a] + 1
            ta = ta + curr_segment_value
            t[a] = ta
            prev_surf_dist[a] = surf_dist
            prev_sdf_grad_c[a] = sdf_grad_c
            keep = (ta <= maxt[a]) & (~intersected)
            act = a[keep]

        if extra_outputs is not None:
            extra_outputs['i'] = it
            extra_outputs['weight_sum'] = weight_sum.clone()

        if self.refine_intersection:
            its_t = self._refine(o, d, its_t, trace_eps)

        inv_weight_sum = 1 / weight_sum
        warp_t = warp_t * inv_weight_sum
        warp_t_d = (-warp_t[:, None] * weight_d_sum + mixed_sum_d) * inv_weight_sum[:, None]
        warp_weight = weight_sum.clamp(0.0, 1.0)
        warp_weight_d = torch.where(((weight_sum > 0.0) & (weight_sum < 1.0))[:, None], weight_d_sum,
                                    torch.zeros_like(weight_d_sum))
        invalid = (weight_sum < 1e-7) | ~intersects_bbox
        invalid = invalid | ~torch.isfinite(warp_t)  # 0/0 when weight_sum == 0 (reference: same branch via <1e-7)
        warp_t = torch.where(invalid, torch.full_like(warp_t, INF), warp_t)
        z3 = torch.zeros_like(warp_t_d)
        warp_t_d = torch.where(invalid[:, None], z3, warp_t_d)
        warp_weight = torch.where(invalid, torch.zeros_like(warp_weight), warp_weight)
        warp_weight_d = torch.where(invalid[:, None], z3, warp_weight_d)
        return its_t, warp_t, warp_t_d, warp_weight, warp_weight_d

    @torch.no_grad()
    def _refine(self, o, d, its_t, trace_eps):
        """python/shapes.py:245-257 / :323-334."""
        its_t = its_t.clone()
        ref = torch.nonzero(torch.isfinite(its_t)).squeeze(-1)
        i = 0
        while ref.numel() > 0 and i < 10:
            tt = its_t[ref]
            min_dist = self._taps(o[ref] + tt[:, None] * d[ref], 0)
            its_t[ref] = tt + min_dist * (10.0 / (10.0 + i))
            keep = (min_dist <= 0) | (min_dist > trace_eps[ref])
            i += 1
            ref = ref[keep]
        return its_t

    # ---- python/shapes.py:290-339: value-only tracer -----------------------------------
    @torch.no_grad()
    def ray_intersect_non_diff(self, o, d, ray_maxt, active, refine=None):
        N = o.shape[0]
        dt = self.dtype
        d = normalize(d)
        bmin, bmax = self.bbox()
        intersects_bbox, mint, maxt, inside_bbox = self._aabb(o, d, bmin, bmax)
        intersects_bbox = intersects_bbox & ((mint > 0) | inside_bbox)
        active = active & intersects_bbox
        maxt = torch.minimum(maxt, ray_maxt)
        its_t = torch.full((N,), INF, dtype=dt)
        t = torch.where(inside_bbox, torch.zeros_like(mint), mint + 1e-5)
        trace_eps = self.trace_eps * maxt.clamp_min(1.0)
        act = torch.nonzero(active).squeeze(-1)
        step = 0
        while act.numel() > 0 and step < self.max_steps:
            step += 1
            a = act
            ta = t[a]
            sdf_value = self._taps(o[a] + ta[:, None] * d[a], 0)
            if self.step_scale != 1.0:
                sdf_value = sdf_value * self.step_scale
            intersected = sdf_value < trace_eps[a]
            its_t[a] = torch.where(intersected, ta, its_t[a])
            curr = torch.where(intersected, torch.zeros_like(sdf_value), sdf_value.abs())
            keep = (ta <= maxt[a]) & (~intersected)
            ta = ta + curr
            t[a] = ta
            keep = keep & (ta <= maxt[a])
            act = a[keep]
        if self.refine_intersection if refine is None else refine:
            its_t = self._refine(o, d, its_t, trace_eps)
        return its_t

    # ---- python/shapes.py:347-366 ----------------------------------------------------
    def compute_surface_interaction(self, o, d, t):
        """Valid (finite t) lanes only.  Returns dict(t, p, n, wi_z); d may carry grad."""
        p = o + t[:, None] * d
        sdf_value, sdf_grad = self.eval_and_grad(p)
        t_diff = sdf_value / dot(sdf_grad, -d).detach()
        t2 = replace_grad(t, t_diff)
        p2 = o + t2[:, None] * d
        n = normalize(self.eval_grad(p2))
        wi_z = dot(n, -d)  # si.to_local(-ray.d).z
        return dict(t=t2, p=p2, n=n, wi_z=wi_z)


# --------------------------------------------------------------------------------------
# Warp field (python/warp.py:7-128)
# --------------------------------------------------------------------------------------


class WarpField2D:
    def __init__(self, sdf, weight_strategy=4, edge_eps=0.05):
        self.max_reparam_depth = -1
        self.sdf = sdf
        self.edge_eps = edge_eps / 4 if weight_strategy == 2 else edge_eps  # python/warp.py:22-23
        self.weight_strategy = weight_strategy
        self.clamping_thresh = 0.0
        self.normalize_warp_field = True

    def weight(self, x, d, sdf_value, sdf_grad, edge_eps):
        """python/warp.py:25-39."""
        bmin, bmax = self.sdf.bbox()
        bbox_dist, bbox_dist_d = bbox_distance_inside_d(x, bmin, bmax)
        use_edge_eps = edge_eps <= bbox_dist
        edge_eps_d = torch.where(use_edge_eps[:, None], torch.zeros_like(bbox_dist_d), bbox_dist_d)
        edge_eps = torch.minimum(edge_eps, bbox_dist)
        inv_edge_eps = 1 / edge_eps
        surf_dist = sdf_value.abs()
        fac = 1 - surf_dist * inv_edge_eps
        w = fac.clamp_min(0.0)
        w_d = -dr_sign(sdf_value)[:, None] * sdf_grad * inv_edge_eps[:, None] + (surf_dist * inv_edge_eps ** 2)[:, None] * edge_eps_d
        w_d = torch.where((fac >= 0.0)[:, None], w_d, torch.zeros_like(w_d))
        edge_eps_d = torch.where(use_edge_eps & (fac >= 0), surf_dist * inv_edge_eps ** 2, torch.zeros_like(surf_dist))
        return w, w_d, edge_eps_d

    def weight_epsilon(self, t):
        """python/warp.py:41-45."""
        if self.weight_strategy == 6:
            return self.edge_eps * t.detach()
        return torch.full_like(t, self.edge_eps)

    def eval(self, x, ray_d, t, dt_dx, active, warp_weight, warp_weight_d, return_active=False):
        """python/warp.py:47-96.  Inputs must be finite on every lane (caller sanitises)."""
        sdf_value, _, sdf_normal, sdf_normal_d, h_mat = self.sdf.eval_all(x)
        h_mat = h_mat.detach()
        if self.normalize_warp_field:
            sdf_normal_d_n, norm_jac = normalize_sqr(sdf_normal_d)
            warp = -sdf_normal_d_n * sdf_value[:, None]
            jac = -(norm_jac @ h_mat) * sdf_value[:, None, None] - outer(sdf_normal_d_n, sdf_normal)
        else:
            sdf_normal_d_n = sdf_normal_d
            warp = -sdf_normal_d_n * sdf_value[:, None]
            jac = -h_mat * sdf_value[:, None, None] - outer(sdf_normal_d_n, sdf_normal)
        xd = x.detach()
        d = ray_d.detach()
        weight, weight_grad, edge_eps_grad = self.weight(xd, d, sdf_value.detach(), sdf_normal.detach(),
                                                         self.weight_epsilon(t))
        weight_grad = weight_grad + edge_eps_grad[:, None] * ray_d * self.edge_eps
        weight_grad = weight_grad * warp_weight[:, None] + weight[:, None] * warp_weight_d
        weight = weight * warp_weight
        weight = weight.detach()
        jac = outer(warp, weight_grad) + weight[:, None, None] * jac
        warp = warp * weight[:, None]
        warp = replace_grad(torch.zeros_like(warp), warp)
        warp = ray_d * torch.clamp_min(t, self.clamping_thresh)[:, None] + warp
        warp = normalize(warp)
        proj_jac = (eye3(ray_d) - outer(ray_d, ray_d)) @ jac
        jac = proj_jac + proj_jac @ outer(ray_d, dt_dx / t[:, None])
        div = jac[:, 0, 0] + jac[:, 1, 1] + jac[:, 2, 2]
        active = active & (weight > 0)
        if return_active:
            return None, active & torch.isfinite(div) & torch.isfinite(warp).all(-1)
        div = torch.where(active, div, torch.zeros_like(div))
        warp = torch.where(active[:, None], warp, ray_d)
        warp = replace_grad(ray_d, warp)
        return warp, div

    def ray_intersect(self, o, d, ray_maxt, depth=0, ray_test=False, reparam=True, active=None):
        """python/warp.py:99-123.  Returns (its_t, d_attached, div, tracer outputs)."""
        N = o.shape[0]
        if active is None:
            active = torch.ones(N, dtype=torch.bool)
        reparam = reparam and ((self.max_reparam_depth < 0) or (depth <= self.max_reparam_depth))
        its_t, warp_t, warp_t_d, warp_weight, warp_weight_d = self.sdf.ray_intersect(
            o.detach(), d.detach(), ray_maxt, active)
        div = torch.ones(N, dtype=o.dtype)
        d_out = d
        if reparam:
            ok = active & torch.isfinite(warp_t)
            sel = torch.nonzero(ok).squeeze(-1)
            if sel.numel() > 0:
                # pass 1 (no grad): which lanes end with weight > 0 (python/warp.py:91); only those carry
                # a derivative, and restricting pass 2 to them keeps 0*inf out of the reverse sweep.
                with torch.no_grad():
                    _, dv0 = self.eval(o[sel] + warp_t[sel, None] * d[sel], d[sel], warp_t[sel], warp_t_d[sel],
                                       torch.ones_like(sel, dtype=torch.bool), warp_weight[sel], warp_weight_d[sel],
                                       return_active=True)
                sel = sel[dv0]
            if sel.numel() > 0:
                ds = d[sel]
                warp, dv = self.eval(o[sel] + warp_t[sel, None] * ds, ds, warp_t[sel], warp_t_d[sel],
                                     torch.ones_like(sel, dtype=torch.bool), warp_weight[sel], warp_weight_d[sel])
                # ray.d = replace_grad(ray.d, warp); div = replace_grad(1, div)  (python/warp.py:114-115)
                d_out = d.index_add(0, sel, warp - warp.detach())
                div = div.index_add(0, sel, dv - dv.detach())
        return its_t, d_out, div, dict(warp_t=warp_t, warp_t_d=warp_t_d, warp_weight=warp_weight,
                                       warp_weight_d=warp_weight_d)


def square_to_von_mises_fisher(u, kappa):
    """mitsuba warp::square_to_von_mises_fisher (Mitsuba 3 include/mitsuba/core/warp.h, the low-distortion variant on the
    concentric disk mapping; Jakob, "Numerically stable sampling of the von Mises Fisher distribution on S^2"):
    p = concentric_disk(u), r2 = |p|^2, sy = max(1 - r2, 1e-6), cos_theta = 1 + log(sy + (1 - sy) exp(-2 kappa)) / kappa,
    direction (p * sqrt((1 - cos_theta^2) / r2), cos_theta).  ref_semantics S19."""
    t = square_to_uniform_disk_concentric(u)
    r2 = (t * t).sum(-1)
    sy = (1 - r2).clamp_min(1e-6)
    cos_theta = 1 + torch.log(sy + (1 - sy) * math.exp(-2 * kappa)) / kappa
    sc = torch.sqrt(((1 - cos_theta * cos_theta) / torch.where(r2 > 0, r2, torch.ones_like(r2))).clamp_min(0.0))
    sc = torch.where(r2 > 0, sc, torch.zeros_like(sc))
    return torch.stack([t[:, 0] * sc, t[:, 1] * sc, cos_theta], -1)


def square_to_von_mises_fisher_pdf(v, kappa):
    """exp(kappa (v.z - 1)) kappa / (2 pi (1 - exp(-2 kappa)))"""
    return torch.exp(kappa * (v[:, 2] - 1)) * (kappa / (2 * math.pi)) / (1 - math.exp(-2 * kappa))


class WarpFieldConvolution:
    """python/warp.py:126-176 + python/warp_conv.py (Bangaru et al. 2020, the comparison baseline): the warp field is a convolution of
    the SDF's own motion over `n_aux_rays` auxiliary rays drawn from a von Mises-Fisher lobe around the ray.

    In the reference the reparameterisation is a dr.CustomOp: its `eval` returns (ray.d, 0) and draws NO samples; the auxiliary rays
    are only traced when the AD graph is traversed (`backward`, python/warp_conv.py:185-282), with samples taken from the integrator's
    sampler in the state it has THEN -- after every forward draw of the lane -- and the CustomOps are visited in reverse creation
    order (BSDF-sampled ray, shadow ray, camera ray).  Restated here with placeholders: `ray_intersect` returns
    d_att = normalize(d + ph_dir), div = 1 + ph_div with zero-valued leaf tensors and records the call; `conv_backward` (called by
    render_backward after the loss has been differentiated w.r.t. the placeholders) replays the records in reverse, draws the
    auxiliary samples, and differentiates  <g_dir, V_theta> + g_div * div_theta  (both zero-valued, python/warp_conv.py:230-236)
    w.r.t. the SDF and -- through the attached ray origin -- the placeholders of earlier calls.  ref_semantics S19."""

    def __init__(self, sdf, n_aux_rays=16):
        self.sdf = sdf
        self.n_aux_rays = n_aux_rays
        self.max_reparam_depth = -1
        self.kappa = 1e5
        self.power = 3.0
        self.pending = []

    def ray_intersect(self, o, d, ray_maxt, depth=0, ray_test=False, reparam=True, active=None, sampler=None, lanes=None, draw_mask=None):
        """lanes: indices of these rays in the sampler's wavefront (None: all of it, in order); draw_mask: the wavefront lanes whose
        sampler advances when the auxiliary samples are drawn (the `active` mask the REFERENCE passes, which for the BSDF-sampled
        ray is `True` for every lane, python/integrators/sdf_direct_reparam.py:95)."""
        N = o.shape[0]
        if active is None:
            active = torch.ones(N, dtype=torch.bool)
        reparam = reparam and ((self.max_reparam_depth < 0) or (depth <= self.max_reparam_depth))
        its_t = self.sdf.ray_intersect_non_diff(o.detach(), d.detach(), ray_maxt, active)
        div = torch.ones(N, dtype=o.dtype)
        d_out = d
        if reparam:
            assert sampler is not None
            ph_dir = torch.zeros(N, 3, dtype=o.dtype, requires_grad=True)
            ph_div = torch.zeros(N, dtype=o.dtype, requires_grad=True)
            act = active[:, None]
            d_out = torch.where(act, normalize(d + ph_dir), d)          # direction = normalize(ray.d + V_theta) (:233); ray.d[active] = new_d
            div = torch.where(active, div + ph_div, div)                # div = replace_grad(1, div) (python/warp.py:150)
            self.pending.append(dict(o=o, d=d.detach(), maxt=ray_maxt, active=active.clone(), sampler=sampler, ph_dir=ph_dir, ph_div=ph_div,
                                     lanes=lanes, draw_mask=draw_mask))
        return its_t, d_out, div, {}

    def _terms(self, op):
        """(V_theta (N, 3), div_theta (N,)) of one recorded call -- python/warp_conv.py:12-68 per auxiliary ray, :226-236 normalisation."""
        sdf, kappa, power = self.sdf, self.kappa, self.power
        o, d, maxt, active, smp = op['o'], op['d'], op['maxt'], op['active'], op['sampler']
        N, dt = o.shape[0], o.dtype
        fs, ft = coordinate_system(d)                                    # mi.Frame3f(detach(ray.d))
        Z = torch.zeros(N, dtype=dt); dZ = torch.zeros(N, 3, dtype=dt)
        V = torch.zeros(N, 3, dtype=dt); div_lhs = torch.zeros(N, dtype=dt)
        od = o.detach()
        lanes, draw_mask = op['lanes'], op['draw_mask']
        for _ in range(self.n_aux_rays):
            if lanes is None:
                u = smp.next_2d(active.numpy() if draw_mask is None else draw_mask)
            else:
                u = smp.next_2d(draw_mask)[lanes]
            off = square_to_von_mises_fisher(u, kappa)
            omega = fs * off[:, 0:1] + ft * off[:, 1:2] + d * off[:, 2:3]
            pdf = square_to_von_mises_fisher_pdf(off, kappa)
            its = sdf.ray_intersect_non_diff(od, omega, maxt, active)
            hit = active & torch.isfinite(its)
            D = torch.exp(kappa - kappa * dot(d, omega)) - 1.0
            B = torch.ones(N, dtype=dt)
            h = torch.nonzero(hit).squeeze(-1)
            Vg = torch.zeros(N, 3, dtype=dt)                             # V_direct - detach(V_direct): zero valued
            if h.numel() > 0:
                p = od[h] + its[h][:, None] * omega[h]                    # detach(si.p)
                val, g = sdf.eval_and_grad(p)
                n_sq, _ = normalize_sqr(g.detach())
                warp = -n_sq * val[:, None]                               # follows the surface along the normal (:41-43)
                Vd = normalize(p + (warp - warp.detach()) - o[h])         # aux_ray.o = ray.o stays attached (:44)
                Vg = Vg.index_add(0, h, Vd - Vd.detach())
                n = normalize(g.detach())
                B = B.index_copy(0, h, dot(n, -omega[h]) ** 2)            # boundary test (:52-53)
            den = D + B
            ok = den > 1e-7
            den_s = torch.where(ok, den, torch.ones_like(den))
            w = torch.where(ok, den_s ** (-power), torch.zeros_like(den)) / pdf
            tmp1 = (D + 1.0) * torch.where(ok, den_s ** (-(power + 1.0)), torch.zeros_like(den)) * kappa * power
            tmp2 = omega - d * dot(d, omega)[:, None]
            dw = (dr_sign(tmp1) * tmp1.abs().clamp_max(1e10) / pdf)[:, None] * tmp2
            w = torch.where(active, w, torch.zeros_like(w))
            dw = torch.where(active[:, None], dw, torch.zeros_like(dw))
            Z = Z + w
            dZ = dZ + dw
            V = V + w[:, None] * Vg
            div_lhs = div_lhs + dot(dw, Vg)
        Z = Z.clamp_min(1e-8)
        V_theta = V / Z[:, None]
        div_theta = (div_lhs - dot(V_theta, dZ)) / Z
        return V_theta, div_theta

    def conv_backward(self, loss, wrt):
        """Gradients of `loss` w.r.t. the tensors `wrt` including everything that flows through the recorded reparameterisations."""
        ops = self.pending
        self.pending = []
        phs = [t for op in ops for t in (op['ph_dir'], op['ph_div'])]
        g = torch.autograd.grad(loss, list(wrt) + phs, allow_unused=True, retain_graph=True)
        zero = lambda t, gi: torch.zeros_like(t) if gi is None else gi
        acc = [zero(t, gi) for t, gi in zip(wrt, g[:len(wrt)])]
        gph = [zero(t, gi) for t, gi in zip(phs, g[len(wrt):])]
        for k in range(len(ops) - 1, -1, -1):
            g_dir, g_div = gph[2 * k], gph[2 * k + 1]
            V_theta, div_theta = self._terms(ops[k])
            s = (g_dir.detach() * V_theta).sum() + (g_div.detach() * div_theta).sum()
            if not s.requires_grad:
                continue
            earlier = phs[:2 * k]
            gk = torch.autograd.grad(s, list(wrt) + earlier, allow_unused=True, retain_graph=True)
            for i in range(len(wrt)):
                if gk[i] is not None:
                    acc[i] = acc[i] + gk[i]
            for j in range(len(earlier)):
                if gk[len(wrt) + j] is not None:
                    gph[j] = gph[j] + gk[len(wrt) + j]
        return acc


class DummyWarpField:
    """python/warp.py:179-196: value-only trace, no reparameterisation."""

    def __init__(self, sdf):
        self.sdf = sdf

    def ray_intersect(self, o, d, ray_maxt, depth=0, ray_test=False, reparam=True, active=None):
        N = o.shape[0]
        if active is None:
            active = torch.ones(N, dtype=torch.bool)
        its_t = self.sdf.ray_intersect_non_diff(o.detach(), d.detach(), ray_maxt, active)
        return its_t, d, torch.ones(N, dtype=o.dtype), {}


# --------------------------------------------------------------------------------------
# Sensor / film / emitter / BSDF restatements
# --------------------------------------------------------------------------------------


def look_at(origin, target, up):
    """mitsuba ScalarTransform4f.look_at: columns (left, new_up, dir, origin)."""
    origin, target, up = (np.asarray(v, dtype=np.float64) for v in (origin, target, up))
    dir_ = target - origin
    dir_ /= np.linalg.norm(dir_)
    left = np.cross(up, dir_)
    left /= np.linalg.norm(left)
    new_up = np.cross(dir_, left)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = left, new_up, dir_, origin
    return m


def regular_camera_origins(n_sensors, radius=2.0, angle_shift=0.0, height_scale=1.0):
    """python/util.py:84-112 with height_steps = int(n_sensors > 1), hemisphere, vary_height."""
    if n_sensors > 1:
        angles = (np.arange(n_sensors) / n_sensors + angle_shift / n_sensors) * 2 * np.pi
        elevation = 1.15 / height_scale + np.sin(angles * n_sensors / 4) * 0.5
        elevation = np.clip(elevation, 0.0, np.pi / 2 + 0.05)
    else:
        # height_steps == 0 -> falls in the `else` branch as well (python/util.py:101-111)
        angles = (np.arange(1) / 1 + angle_shift) * 2 * np.pi
        elevation = 1.15 / height_scale + np.sin(angles * 1 / 4) * 0.5
        elevation = np.clip(elevation, 0.0, np.pi / 2 + 0.05)
    origins = np.stack([np.cos(angles) * np.sin(elevation) * radius, np.cos(elevation) * radius,
                        np.sin(angles) * np.sin(elevation) * radius], -1)
    return origins + np.array([0.5, 0.0, 0.5])


@dataclass
class Sensor:
    """Perspective sensor + hdrfilm(gaussian) as built by python/util.py:115-138."""
    to_world: np.ndarray
    width: int = 128
    height: int = 128
    fov: float = 39.0           # degrees, x axis
    near: float = 1e-2
    far: float = 1e4
    filter_radius: float = 2.0  # gaussian stddev 0.5 -> radius 4*stddev
    filter_stddev: float = 0.5
    sample_border: bool = True

    @property
    def border(self):
        return int(math.ceil(self.filter_radius - 0.5)) if self.sample_border else 0

    def sample_ray(self, pos_px, dtype):
        """PerspectiveCamera::sample_ray_differential; pos in pixel units."""
        W, H = self.width, self.height
        tan = math.tan(math.radians(self.fov) * 0.5)
        aspect = W / H
        sx = pos_px[:, 0] / W
        sy = pos_px[:, 1] / H
        near_p = torch.stack([(1 - 2 * sx) * tan * self.near, (1 - 2 * sy) * tan / aspect * self.near,
                              torch.full_like(sx, self.near)], -1)
        d_l = normalize(near_p)
        R = torch.as_tensor(self.to_world[:3, :3], dtype=dtype)
        c = torch.as_tensor(self.to_world[:3, 3], dtype=dtype)
        d = d_l @ R.T
        inv_z = 1 / d_l[:, 2]
        o = c + d * (self.near * inv_z)[:, None]
        maxt = (self.far - self.near) * inv_z
        return o, d, maxt

    def sample_direction(self, p):
        """PerspectiveCamera::sample_direction: uv (pixels) and importance*inv_dist^2."""
        W, H = self.width, self.height
        dtype = p.dtype
        tan = math.tan(math.radians(self.fov) * 0.5)
        aspect = W / H
        R = torch.as_tensor(self.to_world[:3, :3], dtype=dtype)
        c = torch.as_tensor(self.to_world[:3, 3], dtype=dtype)
        ref = (p - c) @ R
        sx = 0.5 - 0.5 * ref[:, 0] / (ref[:, 2] * tan)
        sy = 0.5 - 0.5 * aspect * ref[:, 1] / (ref[:, 2] * tan)
        uv = torch.stack([sx * W, sy * H], -1)
        dist = torch.linalg.norm(ref, dim=-1)
        inside = (sx >= 0) & (sx <= 1) & (sy >= 0) & (sy <= 1) & (ref[:, 2] >= self.near) & (ref[:, 2] <= self.far)
        # importance(local_d) * inv_dist^2 = norm * inv_ct^3 / dist^2 = norm * dist / z^3
        norm = 1.0 / ((2 * tan) * (2 * tan / aspect))
        w = norm * dist / ref[:, 2] ** 3
        return uv, torch.where(inside, w, torch.zeros_like(w))


def gaussian_filter(x, stddev, radius):
    """mitsuba GaussianFilter::eval: max(0, exp(alpha x^2) - exp(alpha r^2)), alpha=-1/(2 sigma^2)."""
    alpha = -1.0 / (2.0 * stddev * stddev)
    bias = math.exp(alpha * radius * radius)
    return (torch.exp(alpha * x * x) - bias).clamp_min(0.0)


def film_put(sensor, uv, channels):
    """ImageBlock::put (no normalisation) onto a (H+2b, W+2b, C) accumulator."""
    b = sensor.border
    Wb, Hb = sensor.width + 2 * b, sensor.height + 2 * b
    C = channels.shape[1]
    r = sensor.filter_radius
    pos_f = uv + (b - 0.5)
    lo = torch.ceil(pos_f.detach() - r).to(torch.int64)
    acc = torch.zeros(Hb * Wb, C, dtype=uv.dtype)
    n = int(math.ceil(2 * r)) + 1
    for yi in range(n):
        by = lo[:, 1] + yi
        wy = gaussian_filter(by.to(uv.dtype) - pos_f[:, 1], sensor.filter_stddev, r)
        for xi in range(n):
            bx = lo[:, 0] + xi
            wx = gaussian_filter(bx.to(uv.dtype) - pos_f[:, 0], sensor.filter_stddev, r)
            ok = (bx >= 0) & (bx < Wb) & (by >= 0) & (by < Hb) & (bx.to(uv.dtype) <= pos_f[:, 0].detach() + r) & (
                by.to(uv.dtype) <= pos_f[:, 1].detach() + r)
            idx = torch.where(ok, by * Wb + bx, torch.zeros_like(bx))
            val = channels * torch.where(ok, wx * wy, torch.zeros_like(wx))[:, None]
            acc = acc.index_add(0, idx, val)
    return acc.reshape(Hb, Wb, C)


def film_develop(sensor, acc):
    """HDRFilm::develop: rgb / (W == 0 ? 1 : W), crop the border."""
    b = sensor.border
    core = acc[b:b + sensor.height, b:b + sensor.width]
    w = core[..., 3:4]
    w = torch.where(w == 0, torch.ones_like(w), w)
    return core[..., :3] / w


RAY_EPSILON = 2.0 ** -24 * 1500.0
SHADOW_EPSILON = RAY_EPSILON * 10.0


def square_to_uniform_sphere(u):
    z = 1 - 2 * u[:, 1]
    r = torch.sqrt((1 - z * z).clamp_min(0.0))
    phi = 2 * math.pi * u[:, 0]
    return torch.stack([r * torch.cos(phi), r * torch.sin(phi), z], -1)


class EnvMap:
    """Mitsuba's `envmap` emitter (lat-long bitmap, identity to_world), restated -- oracle/ref_semantics.md S18.

    eval: direction d -> uv = (atan2(d.x, -d.z) / 2pi, acos(d.y) / pi); the bitmap is stored with its first column repeated at the end
    (res = (W + 1, H)); uv.x -= 0.5 / (res.x - 1); uv -= floor(uv); uv *= res - 1; bilinear interpolation of the four texels around it.
    sample_direction: Mitsuba importance-samples luminance * sin(theta) with a hierarchical sample warp (`Hierarchical2D`); restated
    here as a PIECEWISE-CONSTANT distribution over the (res.x - 1) x (res.y - 1) bilinear cells (cell weight = mean luminance of its
    four corners * sin(theta) at the cell centre), uniform inside a cell -- the same target density up to the interpolation inside a
    cell, a different sample warp (deliberate deviation, the reference's estimator is unbiased either way).
    pdf over solid angle = cell probability * (res.x - 1)(res.y - 1) / (2 pi^2 sin theta)."""

    def __init__(self, rgb, scale=1.0, to_world=None):
        """to_world: 3x3 rotation of the emitter (Mitsuba: `v = to_world.inverse() * (-si.wi)` before the lookup, sampled directions
        are mapped with to_world); None = identity."""
        self.R = None if to_world is None else torch.as_tensor(to_world, dtype=torch.float64).reshape(3, 3)
        rgb = torch.as_tensor(rgb, dtype=torch.float64) * scale          # (H, W, 3)
        self.data = torch.cat([rgb, rgb[:, :1]], 1).contiguous()       # (H, W + 1, 3)
        self.H, self.W1 = self.data.shape[0], self.data.shape[1]
        lum = (self.data * torch.tensor([0.212671, 0.715160, 0.072169], dtype=torch.float64)).sum(-1)
        cell = 0.25 * (lum[:-1, :-1] + lum[:-1, 1:] + lum[1:, :-1] + lum[1:, 1:])       # (H - 1, W)
        theta = (torch.arange(self.H - 1, dtype=torch.float64) + 0.5) / (self.H - 1) * math.pi
        w = cell * torch.sin(theta)[:, None]
        w = w / w.sum()
        self.cell_prob = w                                               # (H - 1, W)
        self.marg_cdf = torch.cumsum(w.sum(1), 0)                        # rows
        self.cond_cdf = torch.cumsum(w / w.sum(1, keepdim=True).clamp_min(1e-300), 1)

    def _coord(self, d):
        """direction (unit) -> lookup coordinate c in [0, W) x [0, H - 1]."""
        if self.R is not None:
            d = d @ self.R.to(d.dtype)                                   # to_world^-1 d = R^T d  (row vectors: d R)
        u = torch.atan2(d[:, 0], -d[:, 2]) * (0.5 / math.pi)
        v = torch.acos(d[:, 1].clamp(-1.0, 1.0)) * (1.0 / math.pi)
        u = u - 0.5 / (self.W1 - 1)
        u = u - torch.floor(u)
        v = v - torch.floor(v)
        return u * (self.W1 - 1), v * (self.H - 1)

    def eval(self, d):
        """Radiance arriving from direction d (N, 3) (unit vectors; differentiable w.r.t. d)."""
        cx, cy = self._coord(d)
        ix = torch.minimum(torch.floor(cx.detach()), torch.full_like(cx, self.W1 - 2)).long()
        iy = torch.minimum(torch.floor(cy.detach()), torch.full_like(cy, self.H - 2)).long()
        wx1, wy1 = cx - ix, cy - iy
        wx0, wy0 = 1 - wx1, 1 - wy1
        D = self.data.to(d.dtype)
        return (D[iy, ix] * (wx0 * wy0)[:, None] + D[iy, ix + 1] * (wx1 * wy0)[:, None] +
                D[iy + 1, ix] * (wx0 * wy1)[:, None] + D[iy + 1, ix + 1] * (wx1 * wy1)[:, None])

    def pdf(self, d):
        cx, cy = self._coord(d.detach())
        ix = torch.minimum(torch.floor(cx), torch.full_like(cx, self.W1 - 2)).long()
        iy = torch.minimum(torch.floor(cy), torch.full_like(cy, self.H - 2)).long()
        dl = d.detach() if self.R is None else d.detach() @ self.R.to(d.dtype)
        sin_t = torch.sqrt((1 - dl[:, 1] ** 2).clamp_min(0.0))
        p = self.cell_prob.to(d.dtype)[iy, ix] * ((self.W1 - 1) * (self.H - 1)) / (2 * math.pi ** 2 * sin_t.clamp_min(1e-30))
        return torch.where(sin_t > 0, p, torch.zeros_like(p))

    def sample(self, u):
        """u (N, 2) -> (direction, pdf over solid angle)."""
        dt = u.dtype
        iy = torch.searchsorted(self.marg_cdf.to(dt), u[:, 1].contiguous(), right=True).clamp_max(self.H - 2)
        lo = torch.where(iy > 0, self.marg_cdf.to(dt)[(iy - 1).clamp_min(0)], torch.zeros_like(u[:, 1]))
        fy = (u[:, 1] - lo) / (self.marg_cdf.to(dt)[iy] - lo).clamp_min(1e-300)
        cc = self.cond_cdf.to(dt)[iy]                                    # (N, W)
        ix = torch.searchsorted(cc, u[:, 0:1].contiguous(), right=True).squeeze(1).clamp_max(self.W1 - 2)
        lo = torch.where(ix > 0, cc.gather(1, (ix - 1).clamp_min(0)[:, None]).squeeze(1), torch.zeros_like(u[:, 0]))
        hi = cc.gather(1, ix[:, None]).squeeze(1)
        fx = (u[:, 0] - lo) / (hi - lo).clamp_min(1e-300)
        cx, cy = ix.to(dt) + fx.clamp(0, 1), iy.to(dt) + fy.clamp(0, 1)
        uu = cx / (self.W1 - 1) + 0.5 / (self.W1 - 1)
        vv = cy / (self.H - 1)
        theta, phi = vv * math.pi, uu * 2 * math.pi
        st = torch.sin(theta)
        d = torch.stack([torch.sin(phi) * st, torch.cos(theta), -torch.cos(phi) * st], -1)
        if self.R is not None:
            d = d @ self.R.to(dt).T                                      # to_world d_local
        return d, self.pdf(d)


@dataclass
class Scene:
    """Minimal stand-in for the Mitsuba scene the integrators query."""
    sdf: Grid3d
    sensor: Sensor
    env_radiance: tuple = (1.0, 1.0, 1.0)     # `constant` emitter
    albedo: tuple = (0.5, 0.5, 0.5)           # diffuse reflectance (constant)
    albedo_grid: torch.Tensor = None          # optional (Z,Y,X,3) trilinear gridvolume on [0,1]^3
    env_dist: float = 2.0 * math.sqrt(3.0)    # constant emitter: 2 * scene bsphere radius (unit cube + margin)
    hide_emitters: bool = False
    use_mis: bool = False                     # ReparamIntegrator prop (python/integrators/reparam.py:17): BSDF-sampling MIS branch
    detach_indirect_si: bool = False          # sdf_direct_reparam props (python/integrators/sdf_direct_reparam.py:13-14,44-49)
    decouple_reparam: bool = False
    envmap: 'EnvMap' = None                   # `envmap` emitter instead of the constant one (direct integrator only)


def eval_albedo(scene, p):
    """diffuse BSDF reflectance: constant or trilinear gridvolume lookup (clamp)."""
    if scene.albedo_grid is None:
        return torch.as_tensor(scene.albedo, dtype=p.dtype).expand(p.shape[0], 3)
    g = scene.albedo_grid
    Z, Y, X, _ = g.shape
    res = torch.tensor([X, Y, Z], dtype=p.dtype)
    u = p * res - 0.5
    uf = torch.floor(u.detach())
    a = u - uf
    i0 = uf.to(torch.int64)
    out = 0
    flat = g.reshape(-1, 3)
    for dz in (0, 1):
        iz = (i0[:, 2] + dz).clamp(0, Z - 1)
        wz = a[:, 2] if dz else 1 - a[:, 2]
        for dy in (0, 1):
            iy = (i0[:, 1] + dy).clamp(0, Y - 1)
            wy = a[:, 1] if dy else 1 - a[:, 1]
            for dx in (0, 1):
                ix = (i0[:, 0] + dx).clamp(0, X - 1)
                wx = a[:, 0] if dx else 1 - a[:, 0]
                out = out + flat[(iz * Y + iy) * X + ix] * (wx * wy * wz)[:, None]
    return out


# --------------------------------------------------------------------------------------
# Integrators (python/integrators/*.py)
# --------------------------------------------------------------------------------------


def _ray_intersect(scene, warp_field, o, d, maxt, depth, reparam, active=None, sampler=None, lanes=None, draw_mask=None):
    """ReparamIntegrator.ray_intersect without OptiX (python/integrators/reparam.py:201-232).

    With reparam=False only `its_t` is consumed; the reference still runs the differentiable
    tracer when a WarpField2D is attached (python/shapes.py:116) but its `its_t` follows the same
    stepping rule (:235,238 vs :317-320), so the value-only tracer is used here.
    """
    wf = warp_field if warp_field is not None else DummyWarpField(scene.sdf)
    if isinstance(wf, WarpField2D) and not (reparam and (wf.max_reparam_depth < 0 or depth <= wf.max_reparam_depth)):
        wf = DummyWarpField(scene.sdf)
    if isinstance(wf, WarpFieldConvolution):
        return wf.ray_intersect(o, d, maxt, depth=depth, reparam=reparam, active=active, sampler=sampler, lanes=lanes, draw_mask=draw_mask)
    return wf.ray_intersect(o, d, maxt, depth=depth, reparam=reparam, active=active)


def sample_silhouette(scene, warp_field, sampler, o, d, maxt, reparam_mode):
    """python/integrators/sdf_silhouette_reparam.py:16-29 (always reparameterises)."""
    its_t, d_att, det, _ = _ray_intersect(scene, warp_field, o, d, maxt, 0, True, sampler=sampler)
    valid = torch.isfinite(its_t)
    # (the reference also builds `si` here; its gradient does not reach `result`)
    result = torch.where(valid, torch.ones_like(det), torch.zeros_like(det)) * det
    return result[:, None].expand(-1, 3), d_att, det


def sample_simple_shading(scene, warp_field, sampler, o, d, maxt, reparam_mode):
    """python/integrators/sdf_simple_shading_reparam.py:16-28."""
    its_t, d_att, det, _ = _ray_intersect(scene, warp_field, o, d, maxt, 0, True, sampler=sampler)
    h = torch.nonzero(torch.isfinite(its_t)).squeeze(-1)
    result = torch.zeros(o.shape[0], dtype=o.dtype)
    if h.numel() > 0:
        si = scene.sdf.compute_surface_interaction(o[h], d_att[h], its_t[h])
        l = torch.full((3,), 1.0 / math.sqrt(3.0), dtype=o.dtype)
        shade = (si['n'] @ l).clamp_min(0.0)
        result = result.index_add(0, h, shade * det[h])
    return result[:, None].expand(-1, 3), d_att, det


def sample_direct(scene, warp_field, sampler, o, d, maxt, reparam_mode):
    """python/integrators/sdf_direct_reparam.py:16-84,109-111 (NEE branch, use_mis=False)."""
    sdf = scene.sdf
    dtype = o.dtype
    N = o.shape[0]
    reparametrize = reparam_mode
    its_t, d_att, det, _ = _ray_intersect(scene, warp_field, o, d, maxt, 0, reparametrize, sampler=sampler)
    valid = torch.isfinite(its_t)
    Le = torch.as_tensor(scene.env_radiance, dtype=dtype)
    env = scene.envmap
    # result += throughput * si.emitter(scene).eval(si): environment for misses (:24-31); si.wi = -ray.d of the (attached) primary
    # direction (python/shapes.py:362), so an `envmap` emitter is evaluated differentiably w.r.t. the warped direction
    miss_vis = (~valid) & (not scene.hide_emitters)
    Le_miss = Le.expand(N, 3) if env is None else env.eval(d_att)
    result = det[:, None] * torch.where(miss_vis[:, None], Le_miss, torch.zeros(N, 3, dtype=dtype))

    # ---- emitter sampling (:36-84), on the lanes with a valid si ----
    u = sampler.next_2d(valid.numpy())   # next_2d(active_e): only lanes with a valid interaction advance (:37-40)
    h = torch.nonzero(valid).squeeze(-1)
    if h.numel() == 0:
        return result, d_att, det
    si = sdf.compute_surface_interaction(o[h], d_att[h], its_t[h])
    # si_d = compute_surface_interaction(detach(ray), detach(its_t)) has the same values (python/warp.py:121)
    p, n = si['p'], si['n']
    with torch.no_grad():
        if env is None:
            d_e = square_to_uniform_sphere(u[h])
            ds_pdf = torch.full((h.numel(),), 1.0 / (4 * math.pi), dtype=dtype)
        else:                          # scene.sample_emitter_direction -> envmap.sample_direction (:40)
            d_e, ds_pdf = env.sample(u[h])
        ds_p = p.detach() + d_e * scene.env_dist
        # si.spawn_ray_to(ds.p): offset along detached n, maxt shortened
        mag = (1 + p.detach().abs().max(-1).values) * RAY_EPSILON
        sgn = dr_sign(dot(n.detach(), ds_p - p.detach()))
        off = (mag * sgn)[:, None] * n.detach()
    if scene.detach_indirect_si:      # shadow_ray = si_d.spawn_ray_to(ds.p), si_d = detach(si) (:44-45)
        o_s = p.detach() + off
    elif scene.decouple_reparam:      # shadow_ray = si_d0.spawn_ray_to(ds.p): interaction of the detached ray (:46-47, python/warp.py:121)
        si0 = sdf.compute_surface_interaction(o[h].detach(), d[h].detach(), its_t[h])
        o_s = si0['p'] + off
    else:
        o_s = p + off
    dvec = ds_p - o_s
    dist = torch.linalg.norm(dvec, dim=-1)
    d_s = (dvec / dist[:, None]).detach()  # shadow_ray.d = detach(shadow_ray.d) (:51)
    maxt_s = (dist * (1 - SHADOW_EPSILON)).detach()
    pdf_ok = ds_pdf > 0                # active_e &= ds.pdf != 0 (:42)
    dm = np.zeros(N, dtype=bool)
    dm[h.numpy()[pdf_ok.numpy()]] = True
    its_s, d_s_att, det_e, _ = _ray_intersect(scene, warp_field, o_s, d_s, maxt_s, 1, reparametrize, active=pdf_ok, sampler=sampler, lanes=h, draw_mask=dm)
    occluded = torch.isfinite(its_s)
    cos_o = dot(n, d_s_att)           # wo = si.to_local(shadow_ray.d); only z enters the diffuse BSDF
    cos_i = si['wi_z']
    # ds.emitter.eval(si_e) with si_e.wi = -shadow_ray.d (:61-68): for an envmap this follows the warped shadow direction; / ds.pdf (:74)
    Le_s = Le.expand(h.numel(), 3) if env is None else env.eval(d_s_att)
    emitter_val = torch.where(pdf_ok[:, None], Le_s / torch.where(pdf_ok, ds_pdf, torch.ones_like(ds_pdf))[:, None], torch.zeros_like(Le_s))
    rho = eval_albedo(scene, p)
    gate = (~occluded) & (cos_i > 0) & (cos_o > 0) & pdf_ok
    bsdf_val = rho / math.pi * cos_o[:, None]
    nee = bsdf_val * emitter_val
    if scene.use_mis:  # :77-79: power heuristic against the (detached) BSDF pdf of the sampled direction
        bsdf_pdf = torch.where((cos_i > 0) & (cos_o > 0), cos_o.detach() / math.pi, torch.zeros_like(cos_o)).detach()
        nee = nee * mis_weight(ds_pdf, bsdf_pdf)[:, None]
    contrib = det[h, None] * nee * det_e[:, None]
    contrib = torch.where(gate[:, None], contrib, torch.zeros_like(contrib))
    result = result.index_add(0, h, contrib)
    if scene.use_mis:
        # ---- BSDF sampling (python/integrators/sdf_direct_reparam.py:87-107); not built in the CUDA path yet ----
        u1 = sampler.next_1d(valid.numpy())  # noqa: F841  component sample; `active &= si.is_valid()` precedes the draws (:88-91)
        u2 = sampler.next_2d(valid.numpy())
        with torch.no_grad():
            wo_l = square_to_cosine_hemisphere(u2[h])
            bs_pdf = wo_l[:, 2] / math.pi
            act = (cos_i.detach() > 0) & (bs_pdf > 0)               # active &= bs.pdf > 0 (:92); bsdf.sample gates on cos_i
            nd = n.detach()
            fs, ft = sh_frame(nd)
            wo_w = fs * wo_l[:, 0:1] + ft * wo_l[:, 1:2] + nd * wo_l[:, 2:3]   # si_d.to_world(bs.wo)
        o_b = _spawn_offset(p, n, wo_w)                              # si.spawn_ray (attached origin), d detached (:93-94)
        its_b, d_b_att, det_b, _ = _ray_intersect(scene, warp_field, o_b, wo_w, torch.full_like(bs_pdf, INF), 1, reparametrize, active=act,
                                                  sampler=sampler, lanes=h, draw_mask=np.ones(N, dtype=bool))
        escaped = ~torch.isfinite(its_b)
        # bsdf.eval(ctx, si, bs.wo): local direction, so only rho(p) carries a derivative (:97)
        bsdf_val_b = torch.where(act[:, None], rho / math.pi * wo_l[:, 2:3], torch.zeros_like(rho))
        env_pdf = torch.full_like(bs_pdf, 1.0 / (4 * math.pi)) if env is None else env.pdf(wo_w)   # scene.pdf_emitter_direction (:101-103)
        emitter_pdf = torch.where(escaped, env_pdf, torch.zeros_like(bs_pdf))
        w_b = mis_weight(bs_pdf, emitter_pdf)
        inv_pdf = torch.where(bs_pdf > 0, 1 / torch.where(bs_pdf > 0, bs_pdf, torch.ones_like(bs_pdf)), torch.zeros_like(bs_pdf))
        # si_bsdf.emitter(scene).eval(si_bsdf) (:105): si_bsdf.wi = -bsdf_ray.d of the escaped, reparameterised ray
        Le_b = Le.expand(h.numel(), 3) if env is None else env.eval(d_b_att)
        contrib_b = det[h, None] * bsdf_val_b * inv_pdf[:, None] * Le_b * w_b[:, None] * det_b[:, None]
        contrib_b = torch.where((act & escaped)[:, None], contrib_b, torch.zeros_like(contrib_b))
        result = result.index_add(0, h, contrib_b)
    return result, d_att, det


INTEGRATORS = {'sdf_silhouette_reparam': sample_silhouette,
               'sdf_simple_shading_reparam': sample_simple_shading,
               'sdf_direct_reparam': sample_direct}


def render(scene, warp_field, integrator, seed, spp, mode_backward=False, lane_range=None, develop=True, antithetic=False):
    """ReparamIntegrator.render (python/integrators/reparam.py:120-185) + eval_sample (:82-118).

    mode_backward=False == dr.ADMode.Primal (no reparameterisation in the direct integrator).
    lane_range=(lo, hi) restricts the wavefront to global lanes [lo, hi) (multi-GPU partition
    emulation; RNG streams stay keyed by the global lane index).  Returns the developed image
    (H, W, 3) or, with develop=False, the raw accumulator (H+2b, W+2b, 4).
    """
    sensor = scene.sensor
    dtype = scene.sdf.dtype
    b = sensor.border
    Wb, Hb = sensor.width + 2 * b, sensor.height + 2 * b
    wavefront = Wb * Hb * spp
    lo, hi = (0, wavefront) if lane_range is None else lane_range
    idx = torch.arange(lo, hi, dtype=torch.int64)
    pix = idx // spp
    py = pix // Wb
    px = pix - py * Wb
    pos = torch.stack([px - b, py - b], -1).to(dtype)
    sampler = IndependentSampler(seed, wavefront, dtype, lane_offset=lo, lane_count=hi - lo)
    r = sampler.next_2d()
    samples = [(pos + r, sampler)]
    if antithetic:  # python/integrators/reparam.py:171-178: mirrored position, clone of the sampler after the jitter draw
        import copy
        samples.append((pos - r + 1.0, copy.deepcopy(sampler)))
    acc = None
    for position_sample, smp in samples:
        _wavelength = smp.next_1d()  # consumed even in RGB (python/integrators/reparam.py:90)
        o, d, maxt = sensor.sample_ray(position_sample, dtype)
        rgb, d_att, det = INTEGRATORS[integrator](scene, warp_field, smp, o, d, maxt, mode_backward)
        # re-projection (python/integrators/reparam.py:100-106)
        it_p = o + d_att
        uv, rw = sensor.sample_direction(it_p)
        rw_safe = torch.where(rw > 0, rw, torch.ones_like(rw))
        rw1 = torch.where(rw > 0, rw_safe / rw_safe.detach(), torch.ones_like(rw))
        rw1 = replace_grad(torch.ones_like(rw1), rw1)
        rgb = rw1[:, None] * rgb
        wch = replace_grad(torch.ones_like(det), det * rw1)
        a = film_put(sensor, uv, torch.cat([rgb, wch[:, None]], -1))
        acc = a if acc is None else acc + a
    if not develop:
        return acc
    return film_develop(sensor, acc)


def render_backward(scene, warp_field, integrator, grad_in, seed, spp, antithetic=False):
    """ReparamIntegrator.render_backward (python/integrators/reparam.py:187-190).

    Returns (image, grad wrt sdf.data (Z,Y,X), grad wrt sdf.p (3,), grad wrt albedo grid or None).
    """
    sdf = scene.sdf
    data0, p0, alb0 = sdf.data, sdf.p, scene.albedo_grid
    sdf.data = data0.detach().clone().requires_grad_(True)
    sdf.p = p0.detach().clone().requires_grad_(True)
    if alb0 is not None:
        scene.albedo_grid = alb0.detach().clone().requires_grad_(True)
    try:
        if isinstance(warp_field, WarpFieldConvolution):
            warp_field.pending = []
        image = render(scene, warp_field, integrator, seed, spp, mode_backward=True, antithetic=antithetic)
        loss = (image * grad_in).sum()
        wrt = [sdf.data, sdf.p] + ([scene.albedo_grid] if alb0 is not None else [])
        if isinstance(warp_field, WarpFieldConvolution):
            grads = warp_field.conv_backward(loss, wrt)
        else:
            grads = torch.autograd.grad(loss, wrt, allow_unused=True)
        g_data = grads[0] if grads[0] is not None else torch.zeros_like(sdf.data)
        g_p = grads[1] if grads[1] is not None else torch.zeros_like(sdf.p)
        g_alb = None
        if alb0 is not None:
            g_alb = grads[2] if grads[2] is not None else torch.zeros_like(scene.albedo_grid)
        return image.detach(), g_data, g_p, g_alb
    finally:
        sdf.data, sdf.p, scene.albedo_grid = data0, p0, alb0


# --------------------------------------------------------------------------------------
# Path replay backpropagation integrator (python/integrators/sdf_prb_reparam.py), SURVEY §8f#1
# --------------------------------------------------------------------------------------


def coordinate_system(n):
    """mitsuba `coordinate_system` (Duff et al. 2017), used by `initialize_sh_frame` when dp_du == 0
    (python/shapes.py:360: dp_du is still zero there).  Returns (s, t); ref_semantics S14."""
    sign = torch.where(n[:, 2] >= 0, torch.ones_like(n[:, 2]), -torch.ones_like(n[:, 2]))
    a = -1.0 / (sign + n[:, 2])
    b = n[:, 0] * n[:, 1] * a
    s = torch.stack([sign * n[:, 0] * n[:, 0] * a + 1.0, sign * b, -sign * n[:, 0]], -1)
    t = torch.stack([b, sign + n[:, 1] * n[:, 1] * a, -n[:, 1]], -1)
    return s, t


def sh_frame(n):
    """SurfaceInteraction.initialize_sh_frame: s = coordinate_system(n).first, t = cross(n, s)."""
    s, _ = coordinate_system(n)
    t = torch.linalg.cross(n, s)
    return s, t


def square_to_uniform_disk_concentric(u):
    """mitsuba warp::square_to_uniform_disk_concentric (Shirley-Chiu); ref_semantics S15."""
    x = 2 * u[:, 0] - 1
    y = 2 * u[:, 1] - 1
    is_zero = (x == 0) & (y == 0)
    q13 = x.abs() < y.abs()
    r = torch.where(q13, y, x)
    rp = torch.where(q13, x, y)
    r_safe = torch.where(is_zero, torch.ones_like(r), r)
    phi = 0.25 * math.pi * rp / r_safe
    phi = torch.where(q13, 0.5 * math.pi - phi, phi)
    phi = torch.where(is_zero, torch.zeros_like(phi), phi)
    return torch.stack([r * torch.cos(phi), r * torch.sin(phi)], -1)


def square_to_cosine_hemisphere(u):
    p = square_to_uniform_disk_concentric(u)
    z = torch.sqrt((1 - (p * p).sum(-1)).clamp_min(0.0))
    return torch.cat([p, z[:, None]], -1)


def mis_weight(pdf_a, pdf_b):
    """mitsuba.ad.integrators.common.mis_weight: power heuristic, detached, 0 where not finite."""
    a2 = pdf_a * pdf_a
    b2 = pdf_b * pdf_b
    den = a2 + b2
    w = torch.where(den > 0, a2 / torch.where(den > 0, den, torch.ones_like(den)), torch.zeros_like(den))
    return w.detach()


def _spawn_offset(p, n, d):
    """Interaction.offset_p: p + detach(n) * detach(mulsign((1 + max|p|) * RayEpsilon, n . d)) (ref_semantics S6)."""
    pd, nd = p.detach(), n.detach()
    mag = (1 + pd.abs().max(-1).values) * RAY_EPSILON
    return p + (mag * dr_sign(dot(nd, d.detach())))[:, None] * nd


@dataclass
class PrbConfig:
    """Props of SdfPrbReparamIntegrator (python/integrators/sdf_prb_reparam.py:35-48, reparam.py:14)."""
    max_depth: int = 4
    rr_depth: int = 5


def _clone_sampler(sampler):
    import copy
    return copy.deepcopy(sampler)


def sample_prb(scene, warp_field, sampler, o, d, maxt, cfg, mode='primal', dL=None, state_in=None):
    """SdfPrbReparamIntegrator.sample (python/integrators/sdf_prb_reparam.py:242-439), diffuse BSDF + constant
    environment emitter.  mode 'primal' returns (L, valid_ray); mode 'backward' returns the scalar
    sum over vertices of <dL, Lo> whose autograd gradient is what `dr.backward_from(dL * Lo)` (:428) accumulates.
    """
    primal = mode == 'primal'
    sdf = scene.sdf
    dtype = o.dtype
    N = o.shape[0]
    INV_4PI = 1.0 / (4 * math.pi)
    Le_env = torch.as_tensor(scene.env_radiance, dtype=dtype)
    use_warp = isinstance(warp_field, WarpField2D)
    depth = torch.zeros(N, dtype=torch.int64)
    L = torch.zeros(N, 3, dtype=dtype) if primal else state_in.detach().clone()
    beta = torch.ones(N, 3, dtype=dtype)
    eta = torch.ones(N, dtype=dtype)
    mis_em = torch.ones(N, dtype=dtype)
    active = torch.ones(N, dtype=torch.bool)
    o_prev = torch.zeros(N, 3, dtype=dtype)
    d_prev = torch.zeros(N, 3, dtype=dtype)
    t_prev = torch.full((N,), INF, dtype=dtype)
    o_cur, d_cur, maxt_cur = o.detach().clone(), d.detach().clone(), maxt.detach().clone()
    t_cur = sdf.ray_intersect_non_diff(o_cur, d_cur, maxt_cur, active)  # ray_intersect_preliminary (:264)
    valid_ray = torch.isfinite(t_cur) | (not scene.hide_emitters)
    if use_warp:  # :267-272
        m = warp_field.max_reparam_depth
        reparam_max_depth = m if m >= 0 else 100000
    else:
        reparam_max_depth = cfg.max_depth
    shade_gradient_max_depth = reparam_max_depth
    loss = torch.zeros((), dtype=dtype)
    big = torch.full((N,), INF, dtype=dtype)

    for _it in range(cfg.max_depth):  # loop.set_max_iterations(self.max_depth) (:283)
        # the sampler advances on every lane; values of inactive lanes are never used
        u_em = sampler.next_2d()
        u_b1 = sampler.next_1d()  # noqa: F841  (component sample, unused by `diffuse`)
        u_b2 = sampler.next_2d()
        u_rr = sampler.next_1d()
        if not bool(active.any()):
            break
        A = torch.nonzero(active).squeeze(-1)
        nA = A.numel()
        first = depth[A] == 0
        oc, dc, tc = o_cur[A], d_cur[A], t_cur[A]
        hit = torch.isfinite(tc)

        # ---- reparameterise (differential phase only, :290-296) -----------------------------------
        d_rep = dc
        det = torch.ones(nA, dtype=dtype)
        si_prev = None
        nf = torch.nonzero(~first).squeeze(-1)
        if not primal:
            o_rp = oc
            if nf.numel() > 0:
                si_prev = sdf.compute_surface_interaction(o_prev[A][nf], d_prev[A][nf], t_prev[A][nf])
                o_sp = _spawn_offset(si_prev['p'], si_prev['n'], dc[nf])  # si_prev.spawn_ray(ray_cur.d)
                o_rp = oc.index_add(0, nf, o_sp - oc[nf])
                si_prev = {k: v.detach() for k, v in si_prev.items()}  # dr.disable_grad(si_prev) (:296)
            if use_warp:
                rp_active = depth[A] <= reparam_max_depth
                mt = torch.where(first, maxt_cur[A], big[:nA])
                _, d_rep, div, _ = warp_field.ray_intersect(o_rp, dc, mt, depth=0, ray_test=True, reparam=True, active=rp_active)
                det = torch.where(first, torch.ones_like(div), div)  # ray_reparam_det[first_vertex] = 1

        # ---- current interaction, attached in the differential phase (:298-304) ---------------------
        h = torch.nonzero(hit).squeeze(-1)
        nh = h.numel()
        with torch.set_grad_enabled(not primal):
            if nh > 0:
                si = sdf.compute_surface_interaction(oc[h], d_rep[h], tc[h])
                p, n, cos_i = si['p'], si['n'], si['wi_z']
            Le = torch.zeros(nA, 3, dtype=dtype)
            miss = torch.nonzero(~hit).squeeze(-1)
            if miss.numel() > 0:  # emitter.eval(si_cur): the constant environment for escaped rays
                Le = Le.index_add(0, miss, beta[A][miss] * mis_em[A][miss, None] * Le_env)

        # ---- next event estimation (:306-333) -----------------------------------------------------------
        active_next = (depth[A] + 1 < cfg.max_depth) & hit
        Lr_dir = torch.zeros(nA, 3, dtype=dtype)
        rho = None
        if nh > 0:
            act_em = active_next[h]
            with torch.no_grad():
                d_e = square_to_uniform_sphere(u_em[A][h])
                ds_p = p.detach() + d_e * scene.env_dist
                ds_pdf = INV_4PI
            with torch.set_grad_enabled(not primal):
                o_s = _spawn_offset(p, n, ds_p - p.detach())  # si_cur.spawn_ray_to(ds.p)
                dvec = ds_p - o_s
                dist = torch.linalg.norm(dvec, dim=-1)
                d_s = (dvec / dist[:, None]).detach()  # em_ray.d = detach(em_ray.d) (:317)
                maxt_s = (dist * (1 - SHADOW_EPSILON)).detach()
                its_s, d_s_att, det_e, _ = _ray_intersect(scene, warp_field, o_s, d_s, maxt_s, 1, not primal, active=act_em)
                occluded = torch.isfinite(its_s)
                act_em = act_em & ~occluded
                em_weight = Le_env / ds_pdf
                cos_o = dot(n, d_s_att)
                rho = eval_albedo(scene, p)
                gate = (cos_i > 0) & (cos_o > 0) & act_em
                bsdf_value_em = torch.where(gate[:, None], rho / math.pi * cos_o[:, None], torch.zeros_like(rho))
                bsdf_pdf_em = torch.where(gate, cos_o.detach() / math.pi, torch.zeros_like(cos_o)).detach()
                mis_direct = mis_weight(torch.full_like(bsdf_pdf_em, ds_pdf), bsdf_pdf_em)
                lr = beta[A][h] * mis_direct[:, None] * bsdf_value_em * em_weight * det_e[:, None]
                lr = torch.where(act_em[:, None], lr, torch.zeros_like(lr))
                Lr_dir = Lr_dir.index_add(0, h, lr)

        # ---- BSDF sampling (:335-340) -------------------------------------------------------------------
        with torch.no_grad():
            bsdf_weight = torch.zeros(nA, 3, dtype=dtype)
            bs_pdf = torch.zeros(nA, dtype=dtype)
            wo_w = torch.zeros(nA, 3, dtype=dtype)
            o_next = torch.zeros(nA, 3, dtype=dtype)
            if nh > 0:
                wo_l = square_to_cosine_hemisphere(u_b2[A][h])
                pdf_h = wo_l[:, 2] / math.pi
                ok = active_next[h] & (cos_i.detach() > 0) & (pdf_h > 0)
                bsdf_weight[h] = torch.where(ok[:, None], rho.detach(), torch.zeros_like(rho))
                bs_pdf[h] = torch.where(ok, pdf_h, torch.zeros_like(pdf_h))
                nd = n.detach()
                fs, ft = sh_frame(nd)
                wo_w[h] = fs * wo_l[:, 0:1] + ft * wo_l[:, 1:2] + nd * wo_l[:, 2:3]
                o_next[h] = _spawn_offset(p.detach(), nd, wo_w[h])
            beta_A = beta[A] * bsdf_weight
        L_prev = L[A]
        L_A = (L_prev + Le.detach() + Lr_dir.detach()) if primal else (L_prev - Le.detach() - Lr_dir.detach())

        # ---- stopping criterion (:342-349) --------------------------------------------------------------
        with torch.no_grad():
            beta_max = beta_A.max(-1).values
            active_next = active_next & (beta_max != 0)
            rr_prob = torch.minimum(beta_max * eta[A] ** 2, torch.full_like(beta_max, 0.95))
            rr_active = depth[A] >= cfg.rr_depth
            scale = torch.where(rr_active & (rr_prob > 0), 1 / torch.where(rr_prob > 0, rr_prob, torch.ones_like(rr_prob)),
                                torch.ones_like(rr_prob))
            beta_A = beta_A * scale[:, None]
            rr_continue = u_rr[A] < rr_prob
            active_next = active_next & (~rr_active | rr_continue)

            # ---- next surface (:351-363) ---------------------------------------------------------------
            t_next = sdf.ray_intersect_non_diff(o_next, wo_w, big[:nA], active_next)
            next_hit = torch.isfinite(t_next)
            pdf_em = torch.where(next_hit, torch.zeros(nA, dtype=dtype), torch.full((nA,), INV_4PI, dtype=dtype))
            mis_em_A = mis_weight(bs_pdf, pdf_em)

        if not primal:
            sampler_clone = _clone_sampler(sampler)
            u_em_n = sampler_clone.next_2d()[A]
            _ = sampler_clone.next_1d()
            u_b2_n = sampler_clone.next_2d()[A]
            ann = active_next & next_hit & (depth[A] + 2 < cfg.max_depth)  # active_next_next (:367)
            Lr_dir_next = torch.zeros(nA, 3, dtype=dtype)
            bsdf_val_next = torch.zeros(nA, 3, dtype=dtype)
            k = torch.nonzero(next_hit).squeeze(-1)
            if k.numel() > 0:
                with torch.no_grad():
                    si_n = sdf.compute_surface_interaction(o_next[k], wo_w[k], t_next[k])
                    cos_i_n = si_n['wi_z']
                    d_en = square_to_uniform_sphere(u_em_n[k])  # sample_emitter_direction(si_next, ..., test_vis=True):
                    cos_o_n = dot(si_n['n'], d_en)              # the SDF is not scene geometry, so never occluded
                    rho_n = eval_albedo(scene, si_n['p'])
                    g_n = ann[k] & (cos_i_n > 0) & (cos_o_n > 0)
                    bv = torch.where(g_n[:, None], rho_n / math.pi * cos_o_n[:, None], torch.zeros_like(rho_n))
                    bp = torch.where(g_n, cos_o_n / math.pi, torch.zeros_like(cos_o_n))
                    md = mis_weight(torch.full_like(bp, INV_4PI), bp)
                    lrn = beta_A[k] * md[:, None] * bv * (Le_env / INV_4PI)
                    Lr_dir_next[k] = torch.where(ann[k, None], lrn, torch.zeros_like(lrn))
                    wo_nn = square_to_cosine_hemisphere(u_b2_n[k])
                    g_nn = (cos_i_n > 0) & (wo_nn[:, 2] > 0)
                    bsdf_val_next[k] = torch.where(g_nn[:, None], rho_n / math.pi * wo_nn[:, 2:3], torch.zeros_like(rho_n))
            Le_next = torch.where((active_next & ~next_hit)[:, None], beta_A * mis_em_A[:, None] * Le_env,
                                  torch.zeros(nA, 3, dtype=dtype))
            L_next = L_A - Le_next - Lr_dir_next
            extra = Le_next.clone()
            # previous vertex' BSDF, differentiated through si_cur.p only (:387-406)
            if nf.numel() > 0 and nh > 0:
                # lanes that are both non-first and hit: position within h of every such lane
                in_h = torch.full((nA,), -1, dtype=torch.int64)
                in_h[h] = torch.arange(nh)
                in_nf = torch.full((nA,), -1, dtype=torch.int64)
                in_nf[nf] = torch.arange(nf.numel())
                both = torch.nonzero((~first) & hit).squeeze(-1)
                if both.numel() > 0:
                    pc = p[in_h[both]]
                    pp = si_prev['p'][in_nf[both]]
                    npv = si_prev['n'][in_nf[both]]
                    wo_prev = normalize(pc - pp)
                    cos_p = dot(npv, wo_prev)
                    ok_p = (si_prev['wi_z'][in_nf[both]] > 0) & (cos_p.detach() > 0)
                    sel = both[ok_p]
                    if sel.numel() > 0:
                        cp = cos_p[ok_p]
                        ratio = cp / cp.detach()  # rho_prev / pi cancels
                        extra = extra.index_add(0, sel, L_prev[sel] * ratio[:, None])
            # next vertex' BSDF (:399,406): for `diffuse` its value does not depend on wi_next, the term carries
            # no derivative; written out for completeness with the NaN-free reciprocal
            inv_n = torch.where(bsdf_val_next != 0, 1 / torch.where(bsdf_val_next != 0, bsdf_val_next, torch.ones_like(bsdf_val_next)),
                                torch.zeros_like(bsdf_val_next))
            extra = extra + torch.where(next_hit[:, None], L_next * (bsdf_val_next * inv_n), torch.zeros_like(L_next))
            # indirect term through the current BSDF (:407-412)
            Lr_ind = torch.zeros(nA, 3, dtype=dtype)
            if nh > 0:
                cos_n = dot(n, wo_w[h])
                g_c = active_next[h] & (cos_i > 0) & (cos_n > 0)
                bsdf_val = torch.where(g_c[:, None], rho / math.pi * cos_n[:, None], torch.zeros_like(rho))
                bsdf_val_det = bsdf_weight[h] * bs_pdf[h, None]
                inv_det = torch.where(bsdf_val_det != 0, 1 / torch.where(bsdf_val_det != 0, bsdf_val_det, torch.ones_like(bsdf_val_det)),
                                      torch.zeros_like(bsdf_val_det))
                Lr_ind = Lr_ind.index_add(0, h, L_A[h] * replace_grad(torch.ones_like(bsdf_val), inv_det * bsdf_val))
            Lo = (Le + Lr_dir + Lr_ind) * det[:, None] + extra
            keep = depth[A] <= shade_gradient_max_depth
            Lo = torch.where(keep[:, None], Lo, torch.zeros_like(Lo))
            loss = loss + (dL[A] * Lo).sum()

        # ---- advance (:433-438) ------------------------------------------------------------------------------
        with torch.no_grad():
            L[A] = L_A
            beta[A] = beta_A
            mis_em[A] = mis_em_A
            o_prev[A], d_prev[A], t_prev[A] = oc, dc, tc
            o_cur[A], d_cur[A], t_cur[A] = o_next, wo_w, t_next
            maxt_cur[A] = big[:nA]
            depth[A] = depth[A] + hit.to(torch.int64)
            active = active.clone()
            active[A] = active_next
    if primal:
        return L, valid_ray
    return loss


def _camera_rays(scene, seed, spp, lane_range=None):
    """sample_rays up to the sensor ray (python/integrators/sdf_prb_reparam.py:441-497)."""
    sensor = scene.sensor
    dtype = scene.sdf.dtype
    b = sensor.border
    Wb, Hb = sensor.width + 2 * b, sensor.height + 2 * b
    wavefront = Wb * Hb * spp
    lo, hi = (0, wavefront) if lane_range is None else lane_range
    idx = torch.arange(lo, hi, dtype=torch.int64)
    pix = idx // spp
    py = pix // Wb
    px = pix - py * Wb
    pos = torch.stack([px - b, py - b], -1).to(dtype)
    sampler = IndependentSampler(seed, wavefront, dtype, lane_offset=lo, lane_count=hi - lo)
    pos_f = pos + sampler.next_2d()
    _sample1 = sampler.next_1d()  # `sample1` of sample_ray_differential (:492)
    o, d, maxt = sensor.sample_ray(pos_f, dtype)
    return sampler, pos_f, o, d, maxt


def render_prb(scene, warp_field, seed, spp, cfg=None, lane_range=None, develop=True):
    """SdfPrbReparamIntegrator.render (python/integrators/sdf_prb_reparam.py:54-76)."""
    cfg = cfg or PrbConfig()
    with torch.no_grad():
        sampler, pos_f, o, d, maxt = _camera_rays(scene, seed, spp, lane_range)
        L, _valid = sample_prb(scene, warp_field, sampler, o, d, maxt, cfg, 'primal')
        acc = film_put(scene.sensor, pos_f, torch.cat([L, torch.ones_like(L[:, :1])], -1))
        return film_develop(scene.sensor, acc) if develop else acc


def render_backward_prb(scene, warp_field, grad_in, seed, spp, cfg=None):
    """SdfPrbReparamIntegrator.render_backward (python/integrators/sdf_prb_reparam.py:154-240).

    Returns (image of the re-render, grad wrt sdf.data, grad wrt sdf.p, grad wrt albedo grid or None)."""
    cfg = cfg or PrbConfig()
    if isinstance(warp_field, DummyWarpField):
        raise AttributeError("'DummyWarpField' object has no attribute 'reparam'")  # python/warp.py:179-196
    sdf = scene.sdf
    data0, p0, alb0 = sdf.data, sdf.p, scene.albedo_grid
    sdf.data = data0.detach().clone().requires_grad_(True)
    sdf.p = p0.detach().clone().requires_grad_(True)
    if alb0 is not None:
        scene.albedo_grid = alb0.detach().clone().requires_grad_(True)
    try:
        wrt = [sdf.data, sdf.p] + ([scene.albedo_grid] if alb0 is not None else [])
        sampler, pos_f, o, d, maxt = _camera_rays(scene, seed, spp)
        N = o.shape[0]
        det = torch.ones(N, dtype=o.dtype)
        uv = pos_f
        if isinstance(warp_field, WarpField2D):  # sample_rays with reparam (:499-511)
            m = warp_field.max_reparam_depth
            act = torch.ones(N, dtype=torch.bool) if (m < 0 or 0 <= m) else torch.zeros(N, dtype=torch.bool)
            _, d_att, det, _ = warp_field.ray_intersect(o, d, maxt, depth=0, ray_test=True, reparam=True, active=act)
            uv, _rw = scene.sensor.sample_direction(o + d_att)
        with torch.no_grad():
            L, _valid = sample_prb(scene, warp_field, _clone_sampler(sampler), o, d, maxt, cfg, 'primal')
        Ld = L.clone().requires_grad_(True)
        acc = film_put(scene.sensor, uv, torch.cat([Ld * det[:, None], det[:, None]], -1))  # value = L*weight*det, weight = det
        image = film_develop(scene.sensor, acc)
        loss1 = (image * grad_in).sum()
        g1 = torch.autograd.grad(loss1, [Ld] + wrt, allow_unused=True)
        dL = g1[0]
        loss2 = sample_prb(scene, warp_field, sampler, o, d, maxt, cfg, 'backward', dL=dL, state_in=L)
        g2 = torch.autograd.grad(loss2, wrt, allow_unused=True) if loss2.requires_grad else [None] * len(wrt)
        out = []
        for a, b_, ref in zip(g1[1:], g2, wrt):
            g = torch.zeros_like(ref)
            if a is not None:
                g = g + a
            if b_ is not None:
                g = g + b_
            out.append(g)
        return image.detach(), out[0], out[1], (out[2] if alb0 is not None else None)
    finally:
        sdf.data, sdf.p, scene.albedo_grid = data0, p0, alb0


# --------------------------------------------------------------------------------------
# Synthetic stand-in SDFs -- scenes of the reference are a separate download
# --------------------------------------------------------------------------------------


def _grid_points(res):
    """Nodes at linspace(0,1,R) like create_sphere_sdf (python/shapes.py:557-560)."""
    lin = np.linspace(0, 1, res)
    z, y, x = np.meshgrid(lin, lin, lin, indexing='ij')
    return x, y, z


def box_sdf_grid(res, extents=0.49, smoothing=0.01):
    """variables.update_box_sdf (python/variables.py:161-166) with BoxSDF (python/shapes.py:548-550)."""
    lin = np.linspace(-0.5, 0.5, res)
    z, y, x = np.meshgrid(lin, lin, lin, indexing='ij')
    q = np.stack([np.abs(x), np.abs(y), np.abs(z)], -1) - extents
    return np.linalg.norm(np.maximum(q, 0.0), axis=-1) + np.minimum(q.max(-1), 0.0) - smoothing


def make_sdf(kind, res):
    """'sphere' (python/shapes.py:557: c=.5, r=.3), 'knot' (torus-knot tube), 'blobs' (smooth union)."""
    x, y, z = _grid_points(res)
    if kind == 'sphere':
        sdf = np.sqrt((x - 0.5) ** 2 + (y - 0.5) ** 2 + (z - 0.5) ** 2) - 0.3
    elif kind == 'knot':
        # (2,3) torus knot sampled as a polyline; tube radius 0.045; conservative distance
        s = np.linspace(0, 2 * np.pi, 400, endpoint=False)
        R, r = 0.24, 0.10
        cx = 0.5 + (R + r * np.cos(3 * s)) * np.cos(2 * s)
        cy = 0.5 + r * np.sin(3 * s) * 1.3
        cz = 0.5 + (R + r * np.cos(3 * s)) * np.sin(2 * s)
        P = np.stack([cx, cy, cz], -1)
        pts = np.stack([x.ravel(), y.ravel(), z.ravel()], -1)
        best = np.full(pts.shape[0], np.inf)
        for k in range(0, P.shape[0], 8):
            dd = np.linalg.norm(pts[:, None, :] - P[None, k:k + 8, :], axis=-1).min(1)
            best = np.minimum(best, dd)
        sdf = (best - 0.045).reshape(x.shape)
    elif kind == 'blobs':
        rng = np.random.default_rng(0)
        c = rng.uniform(0.25, 0.75, size=(24, 3))
        rad = rng.uniform(0.05, 0.13, size=24)
        k = 0.03
        acc = np.zeros_like(x)
        for ci, ri in zip(c, rad):
            dd = np.sqrt((x - ci[0]) ** 2 + (y - ci[1]) ** 2 + (z - ci[2]) ** 2) - ri
            acc += np.exp(-dd / k)
        sdf = -k * np.log(acc)
    else:
        raise ValueError(kind)
    sdf = np.maximum(sdf, box_sdf_grid(res))  # bbox constraint, python/variables.py:185-187
    return sdf.astype(np.float32)


def make_scene(kind='sphere', res=32, width=32, height=32, n_sensors=6, sensor_idx=0, dtype=torch.float64,
               albedo_grid_res=None):
    origins = regular_camera_origins(n_sensors)
    tw = look_at(origins[sensor_idx], [0.5, 0.5, 0.5], [0, 1, 0])
    sensor = Sensor(to_world=tw, width=width, height=height)
    sdf = Grid3d(torch.from_numpy(make_sdf(kind, res)).to(dtype))
    alb = None
    if albedo_grid_res is not None:
        rng = np.random.default_rng(1)
        alb = torch.from_numpy(rng.uniform(0.2, 0.8, size=(albedo_grid_res,) * 3 + (3,)).astype(np.float32)).to(dtype)
    return Scene(sdf=sdf, sensor=sensor, albedo_grid=alb)
