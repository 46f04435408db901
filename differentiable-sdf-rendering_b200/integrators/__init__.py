from .reparam import ReparamIntegrator, Properties
from .sdf_direct_reparam import SdfDirectReparamIntegrator
from .sdf_prb_reparam import SdfPrbReparamIntegrator
from .sdf_silhouette_reparam import SdfSilhouetteReparamIntegrator
from .sdf_simple_shading_reparam import SdfSimpleShadingReparamIntegrator

_REGISTRY = {c.integrator_name: c for c in (SdfDirectReparamIntegrator, SdfSilhouetteReparamIntegrator,
                                            SdfSimpleShadingReparamIntegrator, SdfPrbReparamIntegrator)}


def create(name, props=None, sdf=None):
    """mi.register_integrator / load_dict({'type': name}) equivalent."""
    return _REGISTRY[name](props, sdf)
