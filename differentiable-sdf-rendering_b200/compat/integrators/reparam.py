from sdfrt.integrators.reparam import *  # noqa: F401,F403
