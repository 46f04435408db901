"""Host-side mirror of python/warp.py: the warp field objects are parameter holders; the
evaluation itself (WarpField2D.eval and its adjoint) runs inside the CUDA kernels."""
from . import _capi


class WarpField2D:
    """python/warp.py:7-23 (same constructor, same attributes)."""
    is_dummy = False

    def __init__(self, sdf, weight_strategy=4, edge_eps=0.05):
        self.max_reparam_depth = -1
        self.sdf = sdf
        self.edge_eps = edge_eps
        self.weight_strategy = weight_strategy
        self.directional = True
        self.clamping_thresh = 0.0
        self.return_aovs = False
        self.normalize_warp_field = True
        if self.weight_strategy == 2:
            self.edge_eps /= 4

    def warp_params(self):
        if self.return_aovs:
            raise NotImplementedError("AOV outputs are not part of the hot path")
        return dict(kind=_capi.WARP_2D, edge_eps=float(self.edge_eps), weight_strategy=int(self.weight_strategy),
                    clamping_thresh=float(self.clamping_thresh), normalize_warp_field=bool(self.normalize_warp_field),
                    max_reparam_depth=int(self.max_reparam_depth))


class WarpFieldConvolution:
    """python/warp.py:126-136 (same constructor, same attributes): the Bangaru et al. 2020 convolution warp the reference compares
    against (`configs.ConvolutionWarp*`).  All of its work happens in the adjoint sweep (`k_backward<CONV>`); sdf_prb_reparam
    does not take it."""
    is_dummy = False

    def __init__(self, sdf, n_aux_rays=16):
        self.sdf = sdf
        self.n_aux_rays = n_aux_rays
        self.max_reparam_depth = -1
        self.kappa = 1e5
        self.power = 3.0
        self.return_aovs = False

    def warp_params(self):
        if self.return_aovs:
            raise NotImplementedError("AOV outputs are not part of the hot path")
        return dict(kind=_capi.WARP_CONV, max_reparam_depth=int(self.max_reparam_depth), n_aux_rays=int(self.n_aux_rays),
                    kappa=float(self.kappa), power=float(self.power))


class DummyWarpField:
    """python/warp.py:179-196."""
    is_dummy = True

    def __init__(self, sdf):
        self.sdf = sdf
        self.return_aovs = False

    def warp_params(self):
        return dict(kind=_capi.WARP_DUMMY)


def warp_params_of(warp_field):
    """integrator.warp_field may be None (python/integrators/reparam.py:216)."""
    return dict(kind=_capi.WARP_NONE) if warp_field is None else warp_field.warp_params()
