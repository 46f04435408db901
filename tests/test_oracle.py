"""Oracle self-checks: published known answers (PCG32), analytic properties of the cubic B-spline,
autograd-vs-finite-difference of the oracle's own building blocks.  CPU only."""
import math

import numpy as np
import pytest
import torch

from oracle import sdf_oracle as O


def test_pcg32_known_answer():
    # pcg-c-basic demo output for pcg32_srandom_r(&rng, 42u, 54u): the published check vector of PCG32 XSH-RR
    rng = O.PCG32(np.array([42], dtype=np.uint64), np.array([54], dtype=np.uint64))
    got = [int(rng.next_u32()[0]) for _ in range(6)]
    assert got == [0xa15c02b7, 0x7b47f409, 0xba1d3330, 0x83d2f293, 0xbfa4784b, 0xcbed606e]


def test_sampler_float_range_and_lane_independence():
    s = O.IndependentSampler(7, 1024, torch.float32)
    a = s.next_2d()
    assert a.min() >= 0 and a.max() < 1
    # a lane's stream depends on (seed, global lane index) only -> a sub-range reproduces the same numbers
    s2 = O.IndependentSampler(7, 1024, torch.float32, lane_offset=100, lane_count=50)
    assert torch.equal(s2.next_2d(), a[100:150])
    assert not torch.equal(O.IndependentSampler(8, 1024, torch.float32).next_2d(), a)


def test_bspline_partition_of_unity_and_derivative_weights():
    a = torch.linspace(0, 1, 33, dtype=torch.float64)
    w, dw, ddw = O._bspline_w(a)
    assert torch.allclose(w.sum(-1), torch.ones_like(a), atol=1e-15)
    assert torch.allclose(dw.sum(-1), torch.zeros_like(a), atol=1e-15)
    assert torch.allclose(ddw.sum(-1), torch.zeros_like(a), atol=1e-15)
    # derivative weights are the derivatives of the value weights
    a2 = a.clone().requires_grad_(True)
    w2, dw2, _ = O._bspline_w(a2)
    for k in range(4):
        g, = torch.autograd.grad(w2[:, k].sum(), a2, retain_graph=True)
        assert torch.allclose(g, dw2[:, k].detach(), atol=1e-14)


def test_cubic_reproduces_linear_fields_and_their_gradient():
    # texel i sits at (i + 0.5) / R; uniform cubic B-splines reproduce linear functions in the interior
    R = 16
    c = (torch.arange(R, dtype=torch.float64) + 0.5) / R
    z, y, x = torch.meshgrid(c, c, c, indexing='ij')
    coef = torch.tensor([0.3, -1.2, 2.0], dtype=torch.float64)
    grid = coef[0] * x + coef[1] * y + coef[2] * z + 0.7
    sdf = O.Grid3d(grid)
    pts = torch.rand(200, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(0)) * 0.6 + 0.2
    v, g, H = sdf._taps(pts, 2)
    assert torch.allclose(v, pts @ coef + 0.7, atol=1e-12)
    assert torch.allclose(g, coef.expand_as(g), atol=1e-10)
    assert H.abs().max() < 1e-8


def test_grad_and_hessian_match_autograd_of_value():
    sc = O.make_scene('blobs', res=16, width=8, height=8)
    sdf = sc.sdf
    pts = (torch.rand(64, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(1)) * 1.1 - 0.05).requires_grad_(True)
    v, g, H = sdf._taps(pts, 2)
    gv, = torch.autograd.grad(v.sum(), pts, create_graph=True)
    assert torch.allclose(gv, g, atol=1e-9)
    for k in range(3):
        hk, = torch.autograd.grad(g[:, k].sum(), pts, retain_graph=True)
        assert torch.allclose(hk, H[:, k, :], atol=1e-7)


def test_sphere_grid_close_to_analytic_sdf():
    sc = O.make_scene('sphere', res=64, width=8, height=8)
    pts = torch.rand(500, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(2)) * 0.5 + 0.25
    v, g = sc.sdf.eval_and_grad(pts)
    exact = torch.linalg.norm(pts - 0.5, dim=-1) - 0.3
    far = (pts - 0.5).norm(dim=-1) > 0.1
    # nodes are sampled at linspace(0,1,R) but looked up at (i+.5)/R (reference behaviour): ~half-texel scale error
    assert (v - exact)[far].abs().max() < 0.03
    assert (g.norm(dim=-1)[far] - 1).abs().max() < 0.1


def test_tracer_hits_sphere_where_expected():
    sc = O.make_scene('sphere', res=64, width=8, height=8)
    o = torch.tensor([[0.5, 0.5, -1.0], [0.5, 0.5, -1.0], [2.0, 2.0, -1.0]], dtype=torch.float64)
    d = torch.tensor([[0.0, 0.0, 1.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]], dtype=torch.float64)
    maxt = torch.full((3,), 1e4, dtype=torch.float64)
    act = torch.ones(3, dtype=torch.bool)
    t_nd = sc.sdf.ray_intersect_non_diff(o, d, maxt, act)
    its_t, warp_t, warp_t_d, ww, wwd = sc.sdf.ray_intersect(o, d, maxt, act)
    assert abs(t_nd[0].item() - 1.2) < 0.02 and math.isinf(t_nd[1].item()) and math.isinf(t_nd[2].item())
    assert torch.allclose(its_t[:1], t_nd[:1], atol=1e-9)
    assert math.isinf(warp_t[2].item()) and ww[2] == 0 and wwd[2].abs().max() == 0  # bbox miss -> invalid (shapes.py:278-283)


def test_film_put_develop_constant_image():
    s = O.Sensor(to_world=np.eye(4), width=8, height=6)
    uv = torch.rand(4000, 2, dtype=torch.float64, generator=torch.Generator().manual_seed(3)) * torch.tensor([12.0, 10.0]) - 2.0
    ch = torch.cat([torch.full((4000, 3), 0.25, dtype=torch.float64), torch.ones(4000, 1, dtype=torch.float64)], -1)
    img = O.film_develop(s, O.film_put(s, uv, ch))
    assert img.shape == (6, 8, 3) and torch.allclose(img, torch.full_like(img, 0.25), atol=1e-12)


def test_sensor_ray_reprojects_to_its_pixel():
    sc = O.make_scene('sphere', res=8, width=20, height=12)
    pos = torch.rand(100, 2, dtype=torch.float64, generator=torch.Generator().manual_seed(4)) * torch.tensor([20.0, 12.0])
    o, d, maxt = sc.sensor.sample_ray(pos, torch.float64)
    assert torch.allclose(d.norm(dim=-1), torch.ones(100, dtype=torch.float64), atol=1e-12)
    uv, w = sc.sensor.sample_direction(o + d)
    assert torch.allclose(uv, pos, atol=1e-9) and (w > 0).all()


@pytest.mark.slow
def test_reparam_gradient_matches_finite_differences_on_translation():
    """The reference's own validation method (figures/result_utils.py:126-161): AD gradient wrt sdf.p vs FD."""
    torch.set_num_threads(max(1, torch.get_num_threads()))
    sc = O.make_scene('sphere', res=24, width=16, height=16)
    wf = O.WarpField2D(sc.sdf, weight_strategy=6, edge_eps=0.01)
    wf.clamping_thresh = 0.05
    g_in = torch.ones(16, 16, 3, dtype=torch.float64)
    gp = torch.stack([O.render_backward(sc, wf, 'sdf_silhouette_reparam', g_in, seed=100 + s, spp=64)[2] for s in range(3)]).mean(0)
    eps, fd = 0.02, []
    for ax in range(1):  # x axis (camera 0 looks along -x: the strongest component)
        vals = []
        for sgn in (+1, -1):
            p = torch.zeros(3, dtype=torch.float64)
            p[ax] = sgn * eps
            sc.sdf.p = p
            with torch.no_grad():
                vals.append(np.mean([O.render(sc, None, 'sdf_silhouette_reparam', seed=200 + s, spp=256).sum().item() for s in range(2)]))
        fd.append((vals[0] - vals[1]) / (2 * eps))
    sc.sdf.p = torch.zeros(3, dtype=torch.float64)
    assert abs(gp[0].item() - fd[0]) < 0.2 * abs(fd[0]), (gp, fd)


@pytest.mark.slow
def test_loss_gradient_matches_finite_differences_uniform_offset():
    """dL/dc for sdf.data += c (sum of the per-voxel gradient) vs the central difference of the primal L1 loss
    (python/shape_opt.py:78-83 with the direct integrator); same method as figures/result_utils.py:126-161."""
    R, W = 32, 32

    def scene(radius, sidx):
        x, y, z = O._grid_points(R)
        sdf = (np.sqrt((x - .5) ** 2 + (y - .5) ** 2 + (z - .5) ** 2) - radius).astype(np.float32)
        sensor = O.Sensor(to_world=O.look_at(O.regular_camera_origins(6)[sidx], [.5, .5, .5], [0, 1, 0]), width=W, height=W)
        return O.Scene(sdf=O.Grid3d(torch.from_numpy(sdf).double()), sensor=sensor)

    views = (0, 1)
    with torch.no_grad():
        targets = [O.render(scene(0.22, i), None, 'sdf_direct_reparam', seed=100 + i, spp=64) for i in views]
    data = scene(0.3, 0).sdf.data.clone()

    def loss_at(d, spp):
        L = 0.0
        for k, i in enumerate(views):
            sc = scene(0.3, i)
            sc.sdf.data = d
            with torch.no_grad():
                L += (O.render(sc, None, 'sdf_direct_reparam', seed=7 + i, spp=spp) - targets[k]).abs().mean().item() / len(views)
        return L

    g_sum = 0.0
    for k, i in enumerate(views):
        sc = scene(0.3, i)
        sc.sdf.data = data
        wf = O.WarpField2D(sc.sdf, weight_strategy=6, edge_eps=0.01)
        wf.clamping_thresh = 0.05
        with torch.no_grad():
            img = O.render(sc, wf, 'sdf_direct_reparam', seed=7 + i, spp=64)
        g_in = torch.sign(img - targets[k]) / img.numel() / len(views)
        g_sum += O.render_backward(sc, wf, 'sdf_direct_reparam', g_in, seed=20 + i, spp=16)[1].sum().item()
    fd = (loss_at(data + 0.01, 128) - loss_at(data - 0.01, 128)) / 0.02
    assert fd < 0 and abs(g_sum - fd) < 0.15 * abs(fd), (g_sum, fd)


def test_use_mis_branch_is_unbiased_against_nee_only():
    """`use_mis` (python/integrators/sdf_direct_reparam.py:77-107, oracle only so far: the CUDA path rejects the prop): emitter
    sampling + BSDF sampling combined with the power heuristic estimate the same image as emitter sampling alone, and the
    backward pass through the extra reparameterised BSDF ray stays finite."""
    sc = O.make_scene('blobs', res=24, width=12, height=12)
    wf = O.WarpField2D(sc.sdf, weight_strategy=6, edge_eps=0.01)
    wf.clamping_thresh = 0.05
    with torch.no_grad():
        a = np.mean([O.render(sc, wf, 'sdf_direct_reparam', seed=s, spp=256).mean().item() for s in range(2)])
        sc.use_mis = True
        b = np.mean([O.render(sc, wf, 'sdf_direct_reparam', seed=10 + s, spp=256).mean().item() for s in range(2)])
    assert abs(a - b) < 5e-3, (a, b)
    g = O.render_backward(sc, wf, 'sdf_direct_reparam', torch.ones(12, 12, 3, dtype=torch.float64), seed=3, spp=8)[1]
    assert torch.isfinite(g).all() and g.abs().max() > 0
