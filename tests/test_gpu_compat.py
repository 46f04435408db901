"""GPU twin of tests/test_reference_drivers.py (-m gpu): the compat modules (`mitsuba`, `drjit`, `variables`, `util`, ... of
sdfrt.compat) on the CUDA context.  The reference's sources do not exist on the GPU box, so the driver code here is a short loop written
for this test that makes the same calls python/shape_opt.py:44-105 makes (mi.load_file, mi.traverse, params.keep, mi.ad.Adam,
scene_config-style variables, mi.render with seed_grad / spp_grad, dr.backward, opt.step, params.update)."""
import importlib
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_compat_modules_drive_the_cuda_path(built):
    import sdfrt
    import sdfrt.compat as compat
    assert torch.cuda.is_available()
    saved_path, saved_mods = list(sys.path), set(sys.modules)
    compat.activate(None)
    try:
        mi = importlib.import_module('mitsuba')
        dr = importlib.import_module('drjit')
        variables = importlib.import_module('variables')
        losses = importlib.import_module('losses')
        reg = importlib.import_module('regularizations')
        util = importlib.import_module('util')
        from oracle import sdf_oracle as O
        target_grid = O.make_sdf('knot', 32)

        def make(grid):
            return lambda integrator='sdf_direct_reparam', **kw: sdfrt.Scene(sdfrt.integrators.create(integrator, sdf=sdfrt.Grid3d(grid)), [])
        compat.register_scene('knotscene', make(target_grid), make(sdfrt.create_sphere_sdf((16, 16, 16), redistance=False)))
        mi.set_variant('cuda_ad_rgb')
        with pytest.raises(RuntimeError):
            mi.set_variant('llvm_ad_rgb')
        sensors = util.get_regular_cameras(4, resx=48, resy=48)
        ref_scene = mi.load_file('/scenes/knotscene/knotscene.xml', integrator='sdf_direct_reparam', sdf_filename='', resx=48, resy=48)
        with dr.suspend_grad():
            refs = [mi.render(ref_scene, sensor=s, seed=41 + i, spp=64) for i, s in enumerate(sensors)]
        scene = mi.load_file('/scenes/knotscene/knotscene.xml', shape_file='dummysdf.xml', integrator='sdf_direct_reparam', resx=48, resy=48)
        sdf_object = scene.integrator().sdf
        wf = sdfrt.WarpField2D(sdf_object, weight_strategy=6, edge_eps=0.01)
        wf.clamping_thresh = 0.05
        scene.integrator().warp_field = wf
        params = mi.traverse(scene)
        assert any('_sdf_' in s.id() for s in scene.shapes())
        params.keep([mi.SDF_KEY])
        opt = mi.ad.Adam(lr=4e-2, params=params, mask_updates=False)
        var = variables.SdfVariable(mi.SDF_KEY, 16, upsample_iter=None, beta=0.5, regularizer=reg.eval_discrete_laplacian_reg,
                                    regularizer_weight=1e-5)
        var.initialize(opt)
        params.update(opt)
        start = opt[mi.SDF_KEY].detach().clone()
        seed, loss_values = 0, []
        for i in range(3):
            loss = mi.Float(0.0)
            for idx, sensor in enumerate(sensors):
                img = mi.render(scene, params=params, sensor=sensor, seed=seed, spp=16, seed_grad=seed + 1 + len(sensors), spp_grad=4)
                seed += 1 + len(sensors)
                view_loss = losses.multiscale_l1(img, refs[idx]) / len(sensors)
                dr.backward(view_loss)
                loss += view_loss
            r = var.eval_regularizer(opt, sdf_object, i)
            if dr.grad_enabled(r):
                dr.backward(r)
            loss += dr.detach(r)
            var.validate_gradient(opt, i)
            loss_values.append(loss[0])
            opt.step()
            var.validate(opt, i)
            var.update_mean(opt, i)
            params.update(opt)
        assert all(np.isfinite(loss_values)) and loss_values[0] > 0
        end = opt[mi.SDF_KEY].detach()
        assert (end - start).abs().max() > 1e-4 and scene.integrator().sdf.data.data_ptr() == end.data_ptr()
        assert loss_values[-1] < loss_values[0]          # three Adam steps towards the target reduce the loss
    finally:
        sys.path[:] = saved_path
        for m in set(sys.modules) - saved_mods:
            del sys.modules[m]
        compat.SCENES.clear()
