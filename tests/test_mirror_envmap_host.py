"""Python mirror plumbing of the `envmap` emitter (Scene.set_envmap -> integrator bind -> context), on the host stand-in of the
context (tests/host_context.py); the arithmetic itself is checked in test_hostmath_parity.py / test_gpu_parity.py."""
import numpy as np
import pytest
import torch


def test_scene_envmap_reaches_the_context_once_per_change(monkeypatch, built):
    import sdfrt
    import host_context
    from oracle import sdf_oracle as O
    host_context.install(monkeypatch)
    W, spp = 12, 2
    osc = O.make_scene('sphere', res=16, width=W, height=W)
    sdf = sdfrt.Grid3d(osc.sdf.data.float().numpy())
    integ = sdfrt.integrators.create('sdf_direct_reparam', {}, sdf)
    wf = sdfrt.WarpField2D(sdf, weight_strategy=6, edge_eps=0.01)
    integ.warp_field = wf
    sensors = [sdfrt.Sensor(osc.sensor.to_world, W, W)] * 2
    scene = sdfrt.Scene(integ, sensors)
    calls = []
    orig = sdf.ctx.set_envmap
    monkeypatch.setattr(sdf.ctx, 'set_envmap', lambda rgb=None, scale=1.0, to_world=None: (calls.append(rgb is not None), orig(rgb, scale, to_world)))
    img_const = sdfrt.render(scene, sensor=0, seed=3, spp=spp)
    assert calls == []                                     # constant emitter: nothing to upload
    bm = np.random.default_rng(0).uniform(0.1, 2.0, size=(6, 12, 3)).astype(np.float32)
    scene.set_envmap(bm)
    img_env = sdfrt.render(scene, sensor=0, seed=3, spp=spp)
    sdfrt.render(scene, sensor=1, seed=4, spp=spp)
    assert calls == [True]                                 # two views, one upload
    assert (img_env - img_const).abs().max() > 1e-2
    osc.envmap = O.EnvMap(bm.astype(np.float64))
    with torch.no_grad():
        img_o = O.render(osc, O.WarpField2D(osc.sdf, weight_strategy=6, edge_eps=0.01), 'sdf_direct_reparam', seed=3, spp=spp)
    assert np.quantile((img_env.double() - img_o).abs().numpy(), 0.98) < 1e-4
    scene.set_envmap(bm * 2.0)
    sdfrt.render(scene, sensor=0, seed=3, spp=spp)
    assert calls == [True, True]
    scene.set_envmap(None)
    img_back = sdfrt.render(scene, sensor=0, seed=3, spp=spp)
    assert calls == [True, True, False] and torch.equal(img_back, img_const)
