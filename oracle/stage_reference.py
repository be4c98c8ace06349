"""Stage the handful of reference files the CPU reference arm needs into the git-ignored `baseline/_ref/` (TEST / BENCH
INFRASTRUCTURE; build container only).

`/root/reference` does not exist on the GPU box, and the reference cannot be pip-installed (no setup.py / pyproject; its
trainer module needs mmcv, yacs, pickle5, ftfy and a CUDA device at import — SURVEY §3.5/§8c).  What CAN travel is the
source of the classes on the hot path: `__graft_entry__.build()` calls `stage()` here, which copies them byte for byte
(never into git history: `baseline/_ref/` is in .gitignore, not in .gpurunignore) so that `bench.py --impl reference`
and `cpu_baseline` time the reference's OWN `DenseCLIP.forward(image, if_test=True)` (AST-extracted and executed
unmodified by oracle/ref_extract.py with LECB_REFERENCE_ROOT pointing at the staged tree) instead of a port.
"""
from __future__ import annotations

import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
MC = os.path.join("project", "my_code")
FILES = [
    os.path.join(MC, "clip", "__init__.py"),
    os.path.join(MC, "clip", "clip.py"),
    os.path.join(MC, "clip", "model.py"),
    os.path.join(MC, "clip", "simple_tokenizer.py"),
    os.path.join(MC, "clip", "bpe_simple_vocab_16e6.txt.gz"),
    os.path.join(MC, "trainers", "Caption_distill_double.py"),
    os.path.join(MC, "trainers", "utils.py"),
    os.path.join(MC, "datasets", "data_helpers.py"),
    os.path.join(MC, "freq_stats.pkl"),
]


def staged_root():
    """The staged tree if it is complete, else None."""
    ok = all(os.path.isfile(os.path.join(DST, f)) for f in FILES)
    return DST if ok else None


def stage(verbose: bool = False):
    """Copy the files when the reference tree is present (build container); a no-op elsewhere."""
    if not os.path.isdir(SRC):
        return staged_root()
    for f in FILES:
        src, dst = os.path.join(SRC, f), os.path.join(DST, f)
        if not os.path.isfile(src):
            raise FileNotFoundError(src)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.isfile(dst) or os.path.getsize(dst) != os.path.getsize(src) or os.path.getmtime(dst) < os.path.getmtime(src):
            shutil.copyfile(src, dst)
            if verbose:
                print("staged", f)
    # tools/bench_reference_gpu.py looks for baseline/_ref/clip/model.py
    os.makedirs(os.path.join(DST, "clip"), exist_ok=True)
    shutil.copyfile(os.path.join(SRC, MC, "clip", "model.py"), os.path.join(DST, "clip", "model.py"))
    return staged_root()


if __name__ == "__main__":
    print(stage(verbose=True))
