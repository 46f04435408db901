"""sdf_prb_reparam (python/integrators/sdf_prb_reparam.py): the autograd oracle's own checks and the
host-compiled kernel math (csrc/sdfrt_prb.cuh) against the oracle and the golden fixtures.  CPU only; the same comparisons
run through the C ABI on the GPU in test_gpu_parity.py."""
import math
import os

import numpy as np
import pytest
import torch

import hostmath as HM
from golden import make_golden as MG
from oracle import sdf_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def make_pair(kind, R, W, albedo_res=None, max_depth=4, rr_depth=5, mrd=-1, warp=True):
    sc = O.make_scene(kind, res=R, width=W, height=W, albedo_grid_res=albedo_res)
    hs = HM.HostScene(sc.sdf.data.numpy(), sc.sensor.to_world, W, W, integrator=3,
                      albedo_grid=None if albedo_res is None else sc.albedo_grid.numpy())
    wf = None
    if warp:
        wf = O.WarpField2D(sc.sdf, weight_strategy=6, edge_eps=0.01)
        wf.clamping_thresh = 0.05
        wf.max_reparam_depth = mrd
    hs.h.warp.kind = 2 if warp else 0
    hs.h.warp.edge_eps, hs.h.warp.weight_strategy, hs.h.warp.clamping_thresh = 0.01, 6, 0.05
    hs.h.warp.normalize_warp_field, hs.h.warp.max_reparam_depth = 1, mrd
    hs.h.prb.max_depth, hs.h.prb.rr_depth = max_depth, rr_depth
    return sc, wf, hs, O.PrbConfig(max_depth=max_depth, rr_depth=rr_depth)


# ---- the oracle's building blocks -------------------------------------------------------------------------------------

def test_sampling_helpers():
    u = torch.from_numpy(np.random.default_rng(0).uniform(size=(4000, 2)))
    w = O.square_to_cosine_hemisphere(u)
    assert torch.allclose(w.norm(dim=-1), torch.ones(4000, dtype=torch.float64), atol=1e-12) and (w[:, 2] >= 0).all()
    assert abs(w[:, 2].mean().item() - 2.0 / 3.0) < 0.01          # E[cos] under the cosine density
    n = O.normalize(torch.from_numpy(np.random.default_rng(1).normal(size=(500, 3))))
    s, t = O.sh_frame(n)
    for a, b in ((s, t), (s, n), (t, n)):
        assert O.dot(a, b).abs().max() < 1e-12
    assert torch.allclose(torch.linalg.cross(s, t), n, atol=1e-12)  # right-handed (s, t, n)
    a, b = torch.tensor([0.5, 0.0, 2.0]), torch.tensor([0.5, 0.0, 0.0])
    assert torch.allclose(O.mis_weight(a, b), torch.tensor([0.5, 0.0, 1.0]))
    z = torch.zeros(1, 2, dtype=torch.float64) + 0.5                # centre of the square -> pole
    assert torch.allclose(O.square_to_cosine_hemisphere(z), torch.tensor([[0.0, 0.0, 1.0]], dtype=torch.float64))


def test_convex_object_radiance_is_albedo_times_environment():
    """A convex diffuse object under a constant environment: the MIS-combined estimator of the reflected radiance has
    expectation rho * Le at every hit (no interreflection), whatever max_depth is."""
    sc = O.make_scene('sphere', res=32, width=16, height=16)
    sampler, pos_f, o, d, maxt = O._camera_rays(sc, seed=5, spp=64)
    with torch.no_grad():
        t = sc.sdf.ray_intersect_non_diff(o, d, maxt, torch.ones(o.shape[0], dtype=torch.bool))
        L, _ = O.sample_prb(sc, None, sampler, o, d, maxt, O.PrbConfig(max_depth=3), 'primal')
    hit = torch.isfinite(t)
    assert hit.sum() > 2000
    assert abs(L[hit].mean().item() - 0.5) < 0.01 and torch.allclose(L[~hit], torch.ones_like(L[~hit]))


def test_depth_one_is_emission_only():
    sc = O.make_scene('sphere', res=24, width=8, height=8)
    img = O.render_prb(sc, None, seed=1, spp=4, cfg=O.PrbConfig(max_depth=1))
    sil = O.render(sc, None, 'sdf_silhouette_reparam', seed=1, spp=4)
    assert torch.allclose(img, 1.0 - sil, atol=1e-12)              # background radiance 1, object black


def test_dummy_warp_field_is_rejected_like_the_reference():
    sc = O.make_scene('sphere', res=16, width=8, height=8)
    with pytest.raises(AttributeError):
        O.render_backward_prb(sc, O.DummyWarpField(sc.sdf), torch.ones(8, 8, 3, dtype=torch.float64), seed=1, spp=1)


@pytest.mark.slow
def test_prb_gradient_matches_finite_differences_uniform_offset():
    """The reference's validation method (figures/result_utils.py:126-161) on an interreflecting scene: sum of the
    per-voxel gradient vs the central difference of the primal render for sdf.data += c."""
    sc = O.make_scene('blobs', res=24, width=16, height=16)
    wf = O.WarpField2D(sc.sdf, weight_strategy=6, edge_eps=0.01)
    wf.clamping_thresh = 0.05
    cfg = O.PrbConfig(max_depth=3)
    g_in = torch.ones(16, 16, 3, dtype=torch.float64)
    ad = np.mean([O.render_backward_prb(sc, wf, g_in, seed=100 + s, spp=64, cfg=cfg)[1].sum().item() for s in range(3)])
    data = sc.sdf.data.clone()

    def total(dv):
        sc.sdf.data = dv
        v = np.mean([O.render_prb(sc, None, seed=200 + s, spp=256, cfg=cfg).sum().item() for s in range(2)])
        sc.sdf.data = data
        return v
    fd = (total(data + 0.01) - total(data - 0.01)) / 0.02
    assert fd > 0 and abs(ad - fd) < 0.1 * abs(fd), (ad, fd)


# ---- host-compiled kernel math vs oracle ------------------------------------------------------------------------------

CASES = [
    ('sphere', 24, None, 3, 5, -1, True),
    ('blobs', 32, 8, 4, 5, -1, True),       # interreflections + reflectance grid
    ('knot', 32, None, 6, 1, -1, True),     # Russian roulette active from the second vertex
    ('blobs', 32, None, 3, 5, 0, True),     # configs.WarpPRBPrimary
    ('blobs', 24, None, 3, 5, -1, False),   # no warp field: shading gradients only
]


@pytest.mark.parametrize('kind,R,albedo_res,max_depth,rr_depth,mrd,warp', CASES)
def test_image_and_gradient_parity(kind, R, albedo_res, max_depth, rr_depth, mrd, warp):
    W, spp = 16, 4
    sc, wf, hs, cfg = make_pair(kind, R, W, albedo_res, max_depth, rr_depth, mrd, warp)
    img_o = O.render_prb(sc, wf, seed=3, spp=spp, cfg=cfg).numpy()
    film, _ = hs.prb_forward(3, spp)
    assert np.abs(hs.develop(film) - img_o).max() < 2e-5
    g_in = MG.grad_image(W)
    img2, g_o, gp_o, ga_o = O.render_backward_prb(sc, wf, torch.from_numpy(g_in), seed=11, spp=spp, cfg=cfg)
    film, L = hs.prb_forward(11, spp)
    assert np.abs(hs.develop(film) - img2.numpy()).max() < 2e-5
    g_h, ga_h, gp_h, st = hs.prb_backward(11, spp, film, L, g_in)
    g_o = g_o.numpy()
    scale = np.abs(g_o).max()
    assert scale > 0 and st[4] > 0
    # stated fp32 tolerance: max-abs <= 2e-2 * ||g||_inf and relative L2 <= 5e-3
    assert np.abs(g_h - g_o).max() < 2e-2 * scale
    assert np.linalg.norm(g_h - g_o) < 5e-3 * np.linalg.norm(g_o)
    assert np.abs(gp_h - gp_o.numpy()).max() < 1e-2 * max(1.0, np.abs(gp_o.numpy()).max())
    if albedo_res is not None:
        ga_o = ga_o.numpy()
        assert np.abs(ga_h - ga_o).max() < 1e-3 * np.abs(ga_o).max()


@pytest.mark.parametrize('name', list(MG.PRB_CASES))
def test_host_math_matches_golden(name):
    kind, R, W, spp, ar, max_depth, rr_depth, mrd = MG.PRB_CASES[name]
    g = dict(np.load(os.path.join(GOLD, name + '.npz')))
    _, _, hs, _ = make_pair(kind, R, W, ar, max_depth, rr_depth, mrd, True)
    film, _ = hs.prb_forward(MG.SEED, spp)
    assert np.abs(hs.develop(film) - g['image']).max() < 2e-5
    film, L = hs.prb_forward(MG.SEED_GRAD, spp)
    assert np.abs(hs.develop(film) - g['image_adj']).max() < 2e-5
    gs, ga, gp, _ = hs.prb_backward(MG.SEED_GRAD, spp, film, L, MG.grad_image(W))
    assert np.abs(gs - g['grad_sdf']).max() < 2e-2 * np.abs(g['grad_sdf']).max()
    assert np.linalg.norm(gs - g['grad_sdf']) < 5e-3 * np.linalg.norm(g['grad_sdf'])
    if ar is not None:
        assert np.abs(ga - g['grad_albedo']).max() < 1e-3 * np.abs(g['grad_albedo']).max()


def test_oracle_reproduces_prb_golden():
    name = 'knot_prb_rr'
    g = dict(np.load(os.path.join(GOLD, name + '.npz')))
    now = MG.run_prb_case(name)
    for k in ('image', 'image_adj', 'grad_sdf', 'grad_p'):
        assert np.allclose(g[k], now[k], rtol=1e-5, atol=1e-6 * max(1.0, np.abs(g[k]).max())), k


def test_rank_partition_adds_up():
    """Film rows split over 2 ranks (interleaved 8-row bands): films and gradients of the ranks add up to the single-rank
    result; sampler streams stay keyed by the global lane index."""
    _, _, hs, _ = make_pair('blobs', 24, 16, None, 3, 5, -1, True)
    spp, g_in = 2, MG.grad_image(16)
    film, L = hs.prb_forward(7, spp)
    g1 = hs.prb_backward(7, spp, film, L, g_in)[0]
    films, parts = [], []
    for r in range(2):
        f, Lr = hs.prb_forward(7, spp, rank=r, world=2)
        films.append(f)
        parts.append((f, Lr))
    fsum = films[0] + films[1]
    assert np.abs(fsum - film).max() < 1e-5
    g2 = sum(hs.prb_backward(7, spp, fsum, Lr, g_in, rank=r, world=2)[0] for r, (_, Lr) in enumerate(parts))
    assert np.abs(g2 - g1).max() < 1e-4 * np.abs(g1).max()


def test_adjoint_is_linear_in_grad_image():
    _, _, hs, _ = make_pair('blobs', 24, 16, None, 3, 5, -1, True)
    film, L = hs.prb_forward(5, 2)
    rng = np.random.default_rng(0)
    a, b = rng.normal(size=(16, 16, 3)), rng.normal(size=(16, 16, 3))
    ga = hs.prb_backward(5, 2, film, L, a)[0]
    gb = hs.prb_backward(5, 2, film, L, b)[0]
    gab = hs.prb_backward(5, 2, film, L, 2 * a - 3 * b)[0]
    assert np.abs(gab - (2 * ga - 3 * gb)).max() < 1e-4 * max(np.abs(ga).max(), np.abs(gb).max())


@pytest.mark.parametrize('kind,R,albedo_res,max_depth,rr_depth,mrd,warp', CASES[1:4])
def test_wavefront_form_equals_per_lane_walker(kind, R, albedo_res, max_depth, rr_depth, mrd, warp):
    """The wavefront form (every segment marched separately, results kept as per-vertex records, the walker re-run from the
    records as segment generator and for the replay) emulated on the host with the per-ray tracers in place of k_march:
    film, L and gradient equal the per-lane walker bit for bit, and no record row is read before it was written (the emulation
    poisons the records with NaN)."""
    _, _, hs, _ = make_pair(kind, R, 16, albedo_res, max_depth, rr_depth, mrd, warp)
    g_in = MG.grad_image(16)
    film, L = hs.prb_forward(11, 4)
    film_w, L_w, _ = hs.prb_wf_forward(11, 4, False)
    assert np.array_equal(film, film_w) and np.array_equal(L, L_w)
    g = hs.prb_backward(11, 4, film, L, g_in)[0]
    film_d, L_d, recs = hs.prb_wf_forward(11, 4, True)
    g_w = hs.prb_wf_backward(11, 4, film_d, L_d, recs, g_in)[0]
    assert np.array_equal(film, film_d) and np.array_equal(g, g_w) and np.isfinite(g_w).all()
