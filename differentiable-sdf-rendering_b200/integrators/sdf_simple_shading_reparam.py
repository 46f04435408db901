from .reparam import ReparamIntegrator


class SdfSimpleShadingReparamIntegrator(ReparamIntegrator):
    """python/integrators/sdf_simple_shading_reparam.py:7-32."""
    integrator_name = 'sdf_simple_shading_reparam'
