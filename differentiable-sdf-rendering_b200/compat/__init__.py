"""Compatibility layer that lets the reference's own drivers -- python/optimize.py, python/shape_opt.py, python/configs.py,
python/opt_configs.py, python/constants.py, UNCHANGED -- run on top of libsdfrt (north star: "so optimize.py and the opt_configs
drive it unchanged").

The reference reaches its runtime through `import mitsuba as mi`, `import drjit as dr` and through its plugin / support modules
(`shapes`, `warp`, `integrators.*`, `variables`, `losses`, `regularizations`, `redistancing`, `util`, `create_video`).  This directory
provides modules of exactly those names: thin adapters onto the sdfrt Python mirror (and through it the C ABI).  Only the names the
drivers touch exist (SURVEY.md 8b lists them); anything else raises AttributeError rather than pretending.

    import sdfrt.compat as compat
    compat.activate('/path/to/differentiable-sdf-rendering/python')       # puts this directory first on sys.path, the reference's last
    compat.register_scene('dragon', make_target_scene, make_optimisable_scene)   # stands in for scenes/dragon/dragon.xml (a download)
    import configs, optimize
    optimize.optimize('dragon', configs.get_config('warp'), 'no-tex-6', out_dir, ref_spp=64)

Scenes: the reference loads Mitsuba XML scenes (a separate download that is not part of the repository).  `mi.load_file(xml, **args)`
is answered from a registry keyed by the scene's name: `sdf_filename == ''` asks for the TARGET scene (python/optimize.py:17-18,
the object that is rendered into the reference images), `shape_file='dummysdf.xml'` for the optimisable one (python/shape_opt.py:44-45).
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
SCENES = {}


def activate(reference_python_dir=None):
    """Make `import mitsuba`, `import drjit`, `import shapes`, ... resolve to this directory (first on sys.path) and the reference's
    driver modules to `reference_python_dir` (last).  A stub `matplotlib` (loss plot of python/shape_opt.py:107-115) is appended after
    everything else, so a real installation wins."""
    for m in ('mitsuba', 'drjit', 'shapes', 'warp', 'integrators', 'variables', 'losses', 'regularizations', 'redistancing', 'util',
              'create_video'):
        if m in sys.modules and not getattr(sys.modules[m], '__file__', '').startswith(_HERE):
            raise RuntimeError(f"module {m!r} was imported before sdfrt.compat.activate()")
    if _HERE not in sys.path:
        sys.path.insert(0, _HERE)
    if reference_python_dir is not None and reference_python_dir not in sys.path:
        sys.path.append(reference_python_dir)
    stubs = os.path.join(_HERE, '_stubs')
    if stubs not in sys.path:
        sys.path.append(stubs)


def register_scene(name, target_factory, optimisable_factory):
    """`name`: the scene name the drivers use (scenes/<name>/<name>.xml).  Factories take the keyword arguments of mi.load_file
    (integrator=, resx=, resy=, spp=, ...) and return an sdfrt.Scene."""
    SCENES[name] = (target_factory, optimisable_factory)
