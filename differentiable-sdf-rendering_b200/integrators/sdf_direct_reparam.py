from .reparam import ReparamIntegrator, Properties


class SdfDirectReparamIntegrator(ReparamIntegrator):
    """python/integrators/sdf_direct_reparam.py:8-114 (emitter-sampling branch)."""
    integrator_name = 'sdf_direct_reparam'

    def __init__(self, props=None, sdf=None):
        super().__init__(props, sdf)
        props = Properties(props or {})
        self.detach_indirect_si = props.get('detach_indirect_si', False)   # :13
        self.decouple_reparam = props.get('decouple_reparam', False)       # :14

    def _bind(self, scene, sensor):
        ctx = super()._bind(scene, sensor)
        # :44-49: which interaction the shadow ray is spawned from (detach_indirect_si wins, as the if/elif there)
        ctx.set_shadow_origin(1 if self.detach_indirect_si else (2 if self.decouple_reparam else 0))
        return ctx
