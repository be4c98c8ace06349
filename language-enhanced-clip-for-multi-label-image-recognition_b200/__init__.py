"""lecb200 — B200-native (sm_100a) dual-prompt CLIP multi-label scoring path.

Drop-in for the hot path of JarvisUSTC/Language-Enhanced-CLIP-For-Multi-label-Image-Recognition
(`DenseCLIP.forward`, trainers/Caption_distill_double.py:401-545): hand-written CUDA kernels behind a
C ABI (include/lecb.h, csrc/liblecb.so), driven from Python through ctypes with torch used only for
device memory, streams and torch.distributed.  No Triton, no dispatch, no CPU fallback."""
from . import _lib  # noqa: F401  (fails loudly when csrc/liblecb.so is missing)
from ._lib import LecbError, launch_count  # noqa: F401

__all__ = ["LecbError", "launch_count"]
