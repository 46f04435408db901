"""The reference's own drivers on top of the boundary (north star: "so optimize.py and the opt_configs drive it unchanged").

`sdfrt.compat` provides modules named like everything the reference imports from its runtime and plugin layer (`mitsuba`, `drjit`,
`shapes`, `warp`, `integrators.*`, `variables`, ...); python/optimize.py, python/shape_opt.py, python/configs.py,
python/opt_configs.py and python/constants.py are imported from /root/reference UNCHANGED and run: reference images of a target
object, then `optimize_shape` for a few iterations of the `no-tex-6` configuration with the `Warp` method.

The reference's sources only exist in the build container (never copied into this repository), where there is no GPU: the compute
behind the mirror is therefore the host stand-in of tests/host_context.py (the kernels' per-lane math compiled for the host).  On a GPU
box the same compat modules sit on the CUDA context (tests/test_gpu_compat.py runs the identical loop through them, with drivers
restated in the test)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

REF = '/root/reference/python'
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference sources are only available in the build container")


@pytest.fixture
def drivers(monkeypatch, built, tmp_path):
    import sdfrt
    import host_context
    host_context.install(monkeypatch)
    import sdfrt.compat as compat
    saved_path, saved_mods = list(sys.path), set(sys.modules)
    compat.activate(REF)
    yield compat
    sys.path[:] = saved_path
    for m in set(sys.modules) - saved_mods:          # the drivers and the compat modules are re-imported by the next test
        del sys.modules[m]
    compat.SCENES.clear()


def _scene_factories(sdfrt):
    from oracle import sdf_oracle as O

    def make(grid):
        def factory(integrator='sdf_direct_reparam', **kw):
            sdf = sdfrt.Grid3d(grid)
            return sdfrt.Scene(sdfrt.integrators.create(integrator, sdf=sdf), [])
        return factory
    target = O.make_sdf('knot', 24)
    return make(target), make(sdfrt.create_sphere_sdf((8, 8, 8), redistance=False))


def test_reference_optimize_runs_unchanged_on_the_compat_layer(drivers, tmp_path, monkeypatch):
    import sdfrt
    drivers.register_scene('knotscene', *_scene_factories(sdfrt))
    configs = importlib.import_module('configs')
    optimize = importlib.import_module('optimize')
    opt_configs = importlib.import_module('opt_configs')
    shape_opt = importlib.import_module('shape_opt')
    for mod in (configs, optimize, opt_configs, shape_opt):
        assert mod.__file__.startswith(REF), mod.__file__          # the reference's files, not copies
    assert importlib.import_module('shapes').__file__.startswith(os.path.dirname(drivers.__file__))

    config = configs.get_config('warp')
    rest = configs.apply_cmdline_args(config, ['--n_iter=3', '--spp=2', '--primal_spp_mult=2'], return_dict=True)
    assert rest == {} and config.n_iter == 3 and config.spp == 2
    monkeypatch.setattr(optimize, 'RENDER_DIR', str(tmp_path / 'renders'))       # the reference tree is read-only
    out = str(tmp_path / 'out')
    optimize.optimize('knotscene', config, 'no-tex-6', out, ref_spp=16, opt_config_args={'resx': '32', 'resy': '32'})

    run_dir = os.path.join(out, 'knotscene', 'no-tex-6', 'warp')
    import json
    meta = json.load(open(os.path.join(run_dir, 'metadata.json')))
    losses = meta['loss_values']
    assert len(losses) == 3 and all(np.isfinite(losses)) and losses[0] > 0
    assert meta['config']['integrator'] == 'sdf_direct_reparam' and meta['opt_config']['resx'] == 32 and meta['opt_config']['batch_size'] == 6
    # 6 reference images (written by render_reference_images, copied by copy_reference_images_to_output_dir), 6 init renders,
    # 3 x 6 optimisation renders, parameter checkpoints of iteration 0, the last iteration and the EMA ("final")
    files = set(os.listdir(run_dir))
    assert {f'ref-{i:02d}.exr' for i in range(6)} <= files and {f'init-{i:02d}.exr' for i in range(6)} <= files
    assert len(os.listdir(os.path.join(run_dir, 'opt'))) == 18
    params = set(os.listdir(os.path.join(run_dir, 'params')))
    assert {'sdf-data-0000.vol', 'sdf-data-0002.vol', 'sdf-data-final.vol'} <= params, params
    # the grid the loop ends with: the 8^3 start of a 64^3 / 2^3 schedule, changed by the three Adam steps, still a re-distanced SDF
    final = sdfrt.volio.read_vol(os.path.join(run_dir, 'params', 'sdf-data-final.vol'))[0]
    start = sdfrt.volio.read_vol(os.path.join(run_dir, 'params', 'sdf-data-0000.vol'))[0]
    assert final.shape[:3] == (8, 8, 8) and np.abs(final - start).max() > 1e-4


def test_reference_optimize_with_the_convolution_warp_baseline(drivers, tmp_path, monkeypatch):
    """The comparison method of the reference's benchmark (`configs.ConvolutionWarp2`, figures/benchmark/benchmark.py:88) through the
    reference's own `optimize.optimize`, unchanged: `configs.get_warpfield` builds our WarpFieldConvolution, the loop runs one iteration
    of `no-tex-6` and the Adam step moves the grid (the gradient came from the auxiliary-ray adjoint, not from a silent fallback)."""
    import sdfrt
    drivers.register_scene('knotscene', *_scene_factories(sdfrt))
    configs = importlib.import_module('configs')
    optimize = importlib.import_module('optimize')
    config = configs.get_config('convolutionwarp2')
    configs.apply_cmdline_args(config, ['--n_iter=1', '--spp=1', '--primal_spp_mult=1'], return_dict=True)
    monkeypatch.setattr(optimize, 'RENDER_DIR', str(tmp_path / 'renders'))
    out = str(tmp_path / 'out')
    seen = []
    orig = sdfrt.WarpFieldConvolution.warp_params
    monkeypatch.setattr(sdfrt.WarpFieldConvolution, 'warp_params', lambda self: (seen.append(self.n_aux_rays), orig(self))[1])
    optimize.optimize('knotscene', config, 'no-tex-6', out, ref_spp=4, opt_config_args={'resx': '16', 'resy': '16'})
    run_dir = os.path.join(out, 'knotscene', 'no-tex-6', config.name)
    import json
    meta = json.load(open(os.path.join(run_dir, 'metadata.json')))
    assert len(meta['loss_values']) == 1 and np.isfinite(meta['loss_values'][0])
    assert seen and set(seen) == {2}                        # every view of the iteration was bound with the convolution warp
    final = sdfrt.volio.read_vol(os.path.join(run_dir, 'params', 'sdf-data-final.vol'))[0]
    start = np.asarray(sdfrt.create_sphere_sdf((8, 8, 8), redistance=False), dtype=np.float32)
    assert np.abs(np.asarray(final, dtype=np.float32).reshape(-1) - start.reshape(-1)).max() > 1e-5


def test_reference_configs_drive_the_mirror_objects(drivers):
    """configs.get_warpfield / opt_configs.SdfConfig build OUR plugin objects with the reference's attribute pokes."""
    import sdfrt
    configs = importlib.import_module('configs')
    opt_configs = importlib.import_module('opt_configs')
    sdf = sdfrt.Grid3d(sdfrt.create_sphere_sdf((8, 8, 8), redistance=False))
    wf = configs.get_config('warp').get_warpfield(sdf)
    assert isinstance(wf, sdfrt.WarpField2D) and (wf.weight_strategy, wf.edge_eps, wf.clamping_thresh) == (6, 0.01, 0.05)
    wp = configs.get_config('warpprimary').get_warpfield(sdf)
    assert wp.max_reparam_depth == 0
    assert isinstance(configs.get_config('onlyshadinggrad').get_warpfield(sdf), sdfrt.DummyWarpField)
    for name, n in (('convolutionwarp', 16), ('convolutionwarp2', 2), ('convolutionwarp32', 32)):      # the Bangaru et al. baseline of figures/benchmark/benchmark.py:86-92
        wc = configs.get_config(name).get_warpfield(sdf)
        assert isinstance(wc, sdfrt.WarpFieldConvolution) and wc.n_aux_rays == n
        assert wc.warp_params() == dict(kind=3, max_reparam_depth=-1, n_aux_rays=n, kappa=1e5, power=3.0)
    cfg, rest = opt_configs.get_opt_config('diffuse-12-hqq', {'sdf_res': '32', 'foo': 'bar'})
    assert rest == {'foo': 'bar'} and isinstance(cfg.variables[0], sdfrt.variables.SdfVariable) and len(cfg.variables) == 2
    assert cfg.variables[1].k == 'main-bsdf.reflectance.volume.data' and len(cfg.sensors) == 12
    idx = [i for i, _ in cfg.get_sensor_iterator(1)]
    assert idx == [(j * 2 + 1) % 12 for j in range(6)]
