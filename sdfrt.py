"""Import alias: the package directory is named `differentiable-sdf-rendering_b200` (not a valid
Python identifier); `import sdfrt` loads it under the name `sdfrt`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'differentiable-sdf-rendering_b200')
_spec = importlib.util.spec_from_file_location('sdfrt', os.path.join(_dir, '__init__.py'), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules['sdfrt'] = _mod
_spec.loader.exec_module(_mod)
