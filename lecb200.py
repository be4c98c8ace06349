"""Import shim: the package directory is named after the reference repository
(`language-enhanced-clip-for-multi-label-image-recognition_b200/`), which is not a valid Python
identifier, so `import lecb200` loads it under this short name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "language-enhanced-clip-for-multi-label-image-recognition_b200")
_spec = importlib.util.spec_from_file_location("lecb200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["lecb200"] = _mod
_spec.loader.exec_module(_mod)
