from .reparam import ReparamIntegrator


class SdfSilhouetteReparamIntegrator(ReparamIntegrator):
    """python/integrators/sdf_silhouette_reparam.py:7-33."""
    integrator_name = 'sdf_silhouette_reparam'
