"""Programmatic dependent launch is only correct if EVERY kernel of the library waits for the previous grid before its first
global-memory access (lecb_common.cuh: pdl_grid_sync / pdl_wait; lecb_host.h: launch_k decides per launch who gets the
attribute, so any kernel may).  This test reads the SASS of the built library — no GPU needed — and checks, per kernel:
  * a griddepcontrol.wait (SASS: ACQBULK) is present;
  * no global load / store / atomic / TMA transfer sits before it in address order (the prologues in front of it only
    initialise barriers, allocate TMEM and prefetch the tensor-map descriptors that arrive as kernel parameters).
One documented exception: stem_conv1_tc_kernel converts its own constant weights (packed at engine construction) while it
waits — loads only, never a store."""
import functools
import re
import shutil
import subprocess

import pytest

from lecb200 import _lib

GLOBAL_OPS = {"LDG", "STG", "LD", "ST", "ATOM", "ATOMG", "RED", "LDGSTS", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP"}
LOAD_OPS = {"LDG", "LD"}


def _opcode(ins):
    """Base opcode of a SASS line ('@!P0 LDG.E.CONSTANT R1, ...' -> 'LDG')."""
    tok = ins.split()
    if tok and tok[0].startswith("@"):
        tok = tok[1:]
    return tok[0].split(".")[0].rstrip(";") if tok else ""
MAY_LOAD_CONSTANTS_FIRST = ("stem_conv1_tc_kernel",)


@functools.lru_cache(maxsize=1)
def _kernels():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    try:
        txt = subprocess.run([exe, "-sass", _lib.LIB_PATH], capture_output=True, text=True, timeout=600).stdout
    except (OSError, subprocess.TimeoutExpired):
        pytest.skip("cuobjdump not available")
    out, name, body = {}, None, []
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                out[name] = body
            name, body = m.group(1), []
        elif name and re.match(r"\s+/\*[0-9a-f]+\*/\s+\S", line):
            body.append(line.split("*/", 1)[1].strip())
    if name:
        out[name] = body
    return out


def test_every_kernel_waits_before_touching_global_memory():
    kernels = _kernels()
    assert len(kernels) > 100, len(kernels)                      # the library holds a few hundred template instantiations
    missing, early = [], []
    for name, body in kernels.items():
        idx = next((i for i, ins in enumerate(body) if "ACQBULK" in ins), None)
        if idx is None:
            missing.append(name)
            continue
        for ins in body[:idx]:
            op = _opcode(ins)
            if op in GLOBAL_OPS:
                if any(k in name for k in MAY_LOAD_CONSTANTS_FIRST) and op in LOAD_OPS:
                    continue
                early.append((name, ins))
                break
    assert not missing, f"kernels without griddepcontrol.wait: {missing[:5]} (+{max(0, len(missing) - 5)} more)"
    assert not early, f"global access before griddepcontrol.wait: {early[:5]}"


def test_persistent_kernels_release_their_dependents():
    """griddepcontrol.launch_dependents (SASS: PREEXIT) is present in every kernel: without it the next kernel's early launch
    degrades to 'when the last CTA exits' (harmless, but the prompt-tuning step loses a third of the gain)."""
    kernels = _kernels()
    missing = [n for n, body in kernels.items() if not any("PREEXIT" in ins for ins in body)]
    assert not missing, missing[:5]
