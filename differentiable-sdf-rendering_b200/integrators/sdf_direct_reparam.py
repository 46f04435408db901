from .reparam import ReparamIntegrator


class SdfDirectReparamIntegrator(ReparamIntegrator):
    """python/integrators/sdf_direct_reparam.py:8-114 (emitter-sampling branch)."""
    integrator_name = 'sdf_direct_reparam'

    def __init__(self, props=None, sdf=None):
        super().__init__(props, sdf)
        props = props or {}
        if props.get('detach_indirect_si', False) or props.get('decouple_reparam', False):
            raise NotImplementedError("detach_indirect_si / decouple_reparam variants are not provided")
