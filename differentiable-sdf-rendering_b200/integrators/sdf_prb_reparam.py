from .reparam import ReparamIntegrator, Properties
from ..warp import DummyWarpField, WarpFieldConvolution


class SdfPrbReparamIntegrator(ReparamIntegrator):
    """python/integrators/sdf_prb_reparam.py:33-521: path replay backpropagation with reparameterised segments
    (emitter sampling + BSDF sampling combined by MIS, Russian roulette from `rr_depth` on).

    Scope of the CUDA path (csrc/sdfrt_prb.cuh): `diffuse` BSDF (constant or grid reflectance) and one `constant`
    environment emitter, i.e. the same shading model as the direct integrator; `WarpField2D` or no warp field.
    """
    integrator_name = 'sdf_prb_reparam'

    def __init__(self, props=None, sdf=None):
        super().__init__(props, sdf)
        props = Properties(props or {})
        self.reparam_max_depth = props.get('reparam_max_depth', self.max_depth)   # :37
        self.shade_gradient_max_depth = self.reparam_max_depth                    # :38
        self.rr_depth = props.get('rr_depth', 5)                                  # :41
        # Bangaru-2020 convolution warp knobs (:43-45): accepted, unused (this integrator takes WarpField2D or no warp field)
        self.reparam_rays = props.get('reparam_rays', 16)
        self.reparam_kappa = props.get('reparam_kappa', 1e5)
        self.reparam_exp = props.get('reparam_exp', 3.0)

    def aovs(self):
        return []

    def _bind(self, scene, sensor):
        if isinstance(self.warp_field, DummyWarpField):
            # the reference calls warp_field.reparam (sdf_prb_reparam.py:181), which DummyWarpField lacks (python/warp.py:179-196)
            raise AttributeError("'DummyWarpField' object has no attribute 'reparam'")
        if isinstance(self.warp_field, WarpFieldConvolution):
            raise NotImplementedError("sdf_prb_reparam with WarpFieldConvolution: the convolution warp is built for the direct-illumination integrators")
        # max_depth = -1: the reference passes the prop through unchanged (python/integrators/reparam.py:14) -- unlike Mitsuba's own
        # prb there is no "-1 means unbounded" conversion.  `loop.set_max_iterations(-1)` wraps to a huge count, but
        # `active_next = (depth + 1 < self.max_depth)` (sdf_prb_reparam.py:307) is then false at the first vertex, so the path ends
        # there: exactly the behaviour of max_depth = 1, which is what is rendered here (oracle/ref_semantics.md S17).
        if self.max_depth == 0 or self.max_depth < -1:
            raise ValueError("max_depth must be >= 1 (or -1)")
        ctx = super()._bind(scene, sensor)
        ctx.set_prb(1 if self.max_depth == -1 else self.max_depth, self.rr_depth)
        if self.warp_field is not None:  # :267-272
            m = self.warp_field.max_reparam_depth
            self.reparam_max_depth = m if m >= 0 else 100000
            self.shade_gradient_max_depth = self.reparam_max_depth
        return ctx
