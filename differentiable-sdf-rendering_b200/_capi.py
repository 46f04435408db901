"""ctypes binding of libsdfrt.so (include/sdfrt.h).  No torch types cross this boundary:
device buffers are passed as integer addresses, streams as void*."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libsdfrt.so')

SDFRT_OK = 0
INTEGRATOR_SILHOUETTE, INTEGRATOR_SIMPLE_SHADING, INTEGRATOR_DIRECT, INTEGRATOR_PRB = 0, 1, 2, 3
WARP_NONE, WARP_DUMMY, WARP_2D, WARP_CONV = 0, 1, 2, 3


class TracerParams(C.Structure):
    _fields_ = [('step_scale', C.c_float), ('trace_eps', C.c_float), ('refine_intersection', C.c_int32),
                ('use_extra_weight', C.c_int32), ('extra_thresh', C.c_float), ('sil_weight_offset', C.c_float),
                ('sil_weight_epsilon', C.c_float), ('weight_power', C.c_int32), ('use_weight_sum_weight', C.c_int32),
                ('max_steps', C.c_int32)]


class WarpParams(C.Structure):
    _fields_ = [('kind', C.c_int32), ('edge_eps', C.c_float), ('weight_strategy', C.c_int32),
                ('clamping_thresh', C.c_float), ('normalize_warp_field', C.c_int32), ('max_reparam_depth', C.c_int32),
                ('n_aux_rays', C.c_int32), ('kappa', C.c_float), ('power', C.c_float)]


class SensorParams(C.Structure):
    _fields_ = [('to_world', C.c_float * 16), ('fov_x_deg', C.c_float), ('near_clip', C.c_float), ('far_clip', C.c_float),
                ('width', C.c_int32), ('height', C.c_int32), ('filter_radius', C.c_float), ('filter_stddev', C.c_float),
                ('sample_border', C.c_int32)]


class ShadingParams(C.Structure):
    _fields_ = [('env_radiance', C.c_float * 3), ('env_dist', C.c_float), ('albedo', C.c_float * 3),
                ('hide_emitters', C.c_int32)]


class PrbParams(C.Structure):
    _fields_ = [('max_depth', C.c_int32), ('rr_depth', C.c_int32)]


class Stats(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ('n_rays', 'n_evals_value', 'n_evals_grad', 'n_evals_hess', 'n_scatter_sites',
                                          'n_albedo_lookups', 'n_steps_max', 'n_hits', 'n_evals_adjoint', 'n_tap_loads', 'n_skipped', 'n_launches', 'n_box_loads')]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class RayOut(C.Structure):
    _fields_ = [('its_t', C.c_void_p), ('warp_t', C.c_void_p), ('warp_t_d', C.c_void_p), ('warp_weight', C.c_void_p),
                ('warp_weight_d', C.c_void_p), ('n_steps', C.c_void_p)]


# every symbol include/sdfrt.h declares: (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    'sdfrt_create': (_P, [C.c_int, C.POINTER(C.c_int)]),
    'sdfrt_destroy': (None, [_P]),
    'sdfrt_last_error': (C.c_char_p, [_P]),
    'sdfrt_version': (C.c_int, []),
    'sdfrt_tracer_defaults': (None, [C.POINTER(TracerParams)]),
    'sdfrt_warp_defaults': (None, [C.POINTER(WarpParams)]),
    'sdfrt_sensor_defaults': (None, [C.POINTER(SensorParams)]),
    'sdfrt_shading_defaults': (None, [C.POINTER(ShadingParams)]),
    'sdfrt_set_sdf': (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    'sdfrt_set_sdf_transform': (C.c_int, [_P, C.POINTER(C.c_float)]),
    'sdfrt_set_albedo': (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int]),
    'sdfrt_set_tracer': (C.c_int, [_P, C.POINTER(TracerParams)]),
    'sdfrt_set_warp': (C.c_int, [_P, C.POINTER(WarpParams)]),
    'sdfrt_set_sensor': (C.c_int, [_P, C.POINTER(SensorParams)]),
    'sdfrt_set_shading': (C.c_int, [_P, C.POINTER(ShadingParams)]),
    'sdfrt_set_integrator': (C.c_int, [_P, C.c_int]),
    'sdfrt_set_antithetic': (C.c_int, [_P, C.c_int]),
    'sdfrt_set_use_mis': (C.c_int, [_P, C.c_int]),
    'sdfrt_set_partition_rows': (C.c_int, [_P, C.c_int, C.c_int, C.c_int]),
    'sdfrt_set_envmap_host': (C.c_int, [_P, C.c_void_p, C.c_int, C.c_int, C.c_float]),
    'sdfrt_set_envmap_transform': (C.c_int, [_P, C.c_void_p]),
    'sdfrt_set_shadow_origin': (C.c_int, [_P, C.c_int]),
    'sdfrt_prb_defaults': (None, [C.POINTER(PrbParams)]),
    'sdfrt_set_prb': (C.c_int, [_P, C.POINTER(PrbParams)]),
    'sdfrt_set_partition': (C.c_int, [_P, C.c_int, C.c_int]),
    'sdfrt_set_lane_window': (C.c_int, [_P, C.c_uint64, C.c_uint64]),
    'sdfrt_set_empty_space_skip': (C.c_int, [_P, C.c_int]),
    'sdfrt_set_view_slot': (C.c_int, [_P, C.c_int]),
    'sdfrt_num_view_slots': (C.c_int, []),
    'sdfrt_adjoint_token': (C.c_uint64, [_P]),
    'sdfrt_redistance_status': (C.c_int, [_P]),
    'sdfrt_film_dims': (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    'sdfrt_wavefront_size': (C.c_uint64, [_P, C.c_int]),
    'sdfrt_render_primal': (C.c_int, [_P, C.c_uint32, C.c_int, _P, _P]),
    'sdfrt_develop': (C.c_int, [_P, _P, _P, _P]),
    'sdfrt_adjoint_accumulate': (C.c_int, [_P, C.c_uint32, C.c_int, _P, _P]),
    'sdfrt_render_view': (C.c_int, [_P, C.c_uint32, C.c_int, _P, C.c_uint32, C.c_int, _P, _P]),
    'sdfrt_adjoint_backprop': (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    'sdfrt_get_stats': (C.c_int, [_P, C.POINTER(Stats), C.c_int, _P]),
    'sdfrt_set_profiling': (C.c_int, [_P, C.c_int]),
    'sdfrt_set_overlap': (C.c_int, [_P, C.c_int]),
    'sdfrt_get_timing': (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    'sdfrt_get_timing_ex': (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_float)]),
    'sdfrt_trace_rays': (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.POINTER(RayOut), _P]),
    'sdfrt_eval_sdf': (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P]),
    'sdfrt_redistance': (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), _P]),
    'sdfrt_adam_step': (C.c_int, [_P, _P, _P, _P, _P, C.c_size_t, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                                  C.c_float, C.c_float, C.c_float, _P]),
    'sdfrt_laplacian_reg': (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P, _P]),
    'sdfrt_render_view_host': (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_uint32, C.c_int,
                                         _P, _P, _P]),
    'sdfrt_render_view_host_ex': (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int,
                                            C.c_uint32, C.c_int, _P, _P, _P, _P]),
}

_lib = None


class SdfrtError(RuntimeError):
    pass


def load():
    """Load libsdfrt.so.  Fails loudly: there is no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SdfrtError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(ctx, rc):
    if rc != SDFRT_OK:
        msg = load().sdfrt_last_error(ctx)
        raise SdfrtError(f"libsdfrt error {rc}: {msg.decode() if msg else '?'}")
