"""Seeded synthetic CLIP weights / inputs — re-exported from the product package's `synth.py`.

Loaded by file path so that the CPU oracle can be used without liblecb.so being built (importing the
`lecb200` package itself requires the CUDA library and fails loudly without it)."""
import importlib.util
import os
import sys

_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                     "language-enhanced-clip-for-multi-label-image-recognition_b200", "synth.py")
_spec = importlib.util.spec_from_file_location("_lecb200_synth", _path)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["_lecb200_synth"] = _mod
_spec.loader.exec_module(_mod)
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})
