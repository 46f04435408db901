"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol include/sdfrt.h
declares; the Python mirror fails loudly (no fallback) when no device is present."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, 'include', 'sdfrt.h')).read()
    return sorted(set(re.findall(r'\b(sdfrt_[a-z_]+)\s*\(', txt)))


def test_header_symbols_are_exported(built):
    import sdfrt
    lib = ctypes.CDLL(sdfrt._capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sdfrt.h but not exported"
    assert set(names) == set(sdfrt._capi.SIGNATURES), set(names) ^ set(sdfrt._capi.SIGNATURES)
    assert lib.sdfrt_version() == 100


def test_library_contains_sm100a_code(built):
    import subprocess
    import sdfrt
    out = subprocess.run(['cuobjdump', '-lelf', sdfrt._capi.LIB_PATH], capture_output=True, text=True).stdout
    assert 'sm_100a' in out


def test_defaults_follow_reference(built):
    import sdfrt
    lib = sdfrt._capi.load()
    t = sdfrt._capi.TracerParams()
    lib.sdfrt_tracer_defaults(ctypes.byref(t))
    # python/shapes.py:28-41
    assert (t.step_scale, t.refine_intersection, t.use_extra_weight, t.weight_power, t.use_weight_sum_weight) == (1.0, 1, 1, 3, 1)
    assert abs(t.trace_eps - 1e-6) < 1e-12 and abs(t.extra_thresh - 0.05) < 1e-8 and abs(t.sil_weight_offset - 0.05) < 1e-8
    w = sdfrt._capi.WarpParams()
    lib.sdfrt_warp_defaults(ctypes.byref(w))
    # python/configs.py:12-40
    assert (w.kind, w.weight_strategy, w.max_reparam_depth, w.normalize_warp_field) == (2, 6, -1, 1)
    assert abs(w.edge_eps - 0.01) < 1e-8 and abs(w.clamping_thresh - 0.05) < 1e-8


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback(built):
    import sdfrt
    st = ctypes.c_int(0)
    assert not sdfrt._capi.load().sdfrt_create(0, ctypes.byref(st)) and st.value < 0
    with pytest.raises(sdfrt.SdfrtError):
        sdfrt.Context(0)
    with pytest.raises(sdfrt.SdfrtError):
        sdfrt.Grid3d(sdfrt.create_sphere_sdf((8, 8, 8), redistance=False))


def test_facade_mirrors_reference_defaults():
    import sdfrt
    wf = sdfrt.WarpField2D(None)
    assert (wf.weight_strategy, wf.edge_eps, wf.clamping_thresh, wf.max_reparam_depth, wf.normalize_warp_field) == (4, 0.05, 0.0, -1, True)
    assert sdfrt.WarpField2D(None, weight_strategy=2, edge_eps=0.04).edge_eps == 0.01   # python/warp.py:22-23
    cams = sdfrt.get_regular_cameras(6)
    assert len(cams) == 6 and cams[0].film().crop_size() == (128, 128)
    from oracle import sdf_oracle as O
    import numpy as np
    assert np.allclose(cams[0].to_world, O.look_at(O.regular_camera_origins(6)[0], [0.5, 0.5, 0.5], [0, 1, 0]))
    a = sdfrt.synthetic.make_sdf('blobs', 16)
    assert np.array_equal(a, O.make_sdf('blobs', 16))


def test_header_is_plain_c(tmp_path):
    """include/sdfrt.h compiles as C99 (no C++-isms, no torch / CUDA types in the signatures) and a C caller links against
    the library's symbols -- the binding a maintainer of a C host would write."""
    import subprocess
    src = tmp_path / 'probe.c'
    src.write_text(
        '#include "sdfrt.h"\n'
        'int probe(void) {\n'
        '    sdfrt_tracer_params t; sdfrt_warp_params w; sdfrt_sensor s; sdfrt_shading sh; sdfrt_prb_params p;\n'
        '    sdfrt_tracer_defaults(&t); sdfrt_warp_defaults(&w); sdfrt_sensor_defaults(&s); sdfrt_shading_defaults(&sh);\n'
        '    sdfrt_prb_defaults(&p);\n'
        '    return sdfrt_version() + p.max_depth + (int)sizeof(sdfrt_stats) + (int)sizeof(sdfrt_ray_out);\n'
        '}\n')
    subprocess.check_call(['gcc', '-std=c99', '-pedantic', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'), '-c', str(src),
                           '-o', str(tmp_path / 'probe.o')])
