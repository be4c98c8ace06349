"""Run the reference's OWN classes on CPU (TEST INFRASTRUCTURE; build-container only).

`/root/reference` exists only in the build container, never on the GPU box, so nothing here may be
imported by `-m gpu` tests, `smoke()` or `bench.py`.  It is used by `oracle/make_golden.py` (to write
`tests/golden/*`) and by the `not gpu` tests that pin `oracle/restatement.py` against the live
reference when the tree is present.

The trainer module cannot be imported: it needs mmcv / pickle5 / dassl / yacs and
`pickle.load(...).cuda()`s a 220k-row file at import (Caption_distill_double.py:35-36).  So the
`ClassDef`s `TextEncoder`, `PromptLearner`, `DenseCLIP` (Caption_distill_double.py:72-101, 104-308,
354-559) are AST-extracted and exec'd *unmodified* in a namespace that provides the names they use;
the loss functions are extracted the same way from trainers/utils.py:85-190 (whose module import
fails on `from numpy import deprecate`).  `clip/model.py` imports directly.

Sanctioned deviation 1: Caption_distill_double.py:447 hard-codes `view(-1, topk, 1024)`,
valid for RN50 only; `embed_dim != 1024` rewrites that one literal to the model's embed dim.
"""
from __future__ import annotations

import ast
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("LECB_REFERENCE_ROOT", "/root/reference")
MC = os.path.join(REF_ROOT, "project", "my_code")


def available() -> bool:
    return os.path.isfile(os.path.join(MC, "trainers", "Caption_distill_double.py"))


def _load_file_module(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_CACHE = {}


def clip_model_module():
    """The reference `clip/model.py` as a module (imports cleanly)."""
    if "model" not in _CACHE:
        _CACHE["model"] = _load_file_module("_lecb_ref_clip_model", os.path.join(MC, "clip", "model.py"))
    return _CACHE["model"]


def clip_package():
    """The reference `clip` package (clip.py + simple_tokenizer.py) with an `ftfy` shim:
    simple_tokenizer.py:6 imports ftfy only for `fix_text`, an identity on the ASCII prompts used here."""
    if "clip" not in _CACHE:
        if "ftfy" not in sys.modules:
            shim = types.ModuleType("ftfy")
            shim.fix_text = lambda s: s
            sys.modules["ftfy"] = shim
        pkg_dir = os.path.join(MC, "clip")
        spec = importlib.util.spec_from_file_location(
            "_lecb_ref_clip", os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir])
        pkg = importlib.util.module_from_spec(spec)
        sys.modules["_lecb_ref_clip"] = pkg
        spec.loader.exec_module(pkg)
        _CACHE["clip"] = pkg
    return _CACHE["clip"]


def _extract(path, names, rewrite=None):
    with open(path, "r") as f:
        tree = ast.parse(f.read())
    picked = [n for n in tree.body if isinstance(n, (ast.ClassDef, ast.FunctionDef)) and n.name in names]
    missing = set(names) - {n.name for n in picked}
    if missing:
        raise RuntimeError(f"reference symbols not found in {path}: {sorted(missing)}")
    mod = ast.Module(body=picked, type_ignores=[])
    if rewrite is not None:
        mod = rewrite.visit(mod)
    ast.fix_missing_locations(mod)
    return compile(mod, path, "exec")


class _EmbedDimLiteral(ast.NodeTransformer):
    def __init__(self, dim):
        self.dim = dim
        self.hits = 0

    def visit_Constant(self, node):
        if node.value == 1024 and isinstance(node.value, int):
            self.hits += 1
            return ast.copy_location(ast.Constant(self.dim), node)
        return node


def trainer_classes(caption_text_feats, embed_dim=1024):
    """-> namespace dict with the reference `TextEncoder`, `PromptLearner`, `DenseCLIP`.

    `caption_text_feats` plays the module-level global of Caption_distill_double.py:35-36."""
    import torch
    import torch.nn as nn
    from torch.nn import functional as F
    from torchvision.models._utils import IntermediateLayerGetter

    clip_pkg = clip_package()
    rewrite = None
    if embed_dim != 1024:
        rewrite = _EmbedDimLiteral(embed_dim)
    code = _extract(os.path.join(MC, "trainers", "Caption_distill_double.py"),
                    ["TextEncoder", "PromptLearner", "DenseCLIP"], rewrite)
    if rewrite is not None and rewrite.hits != 1:
        raise RuntimeError(f"expected exactly one 1024 literal (T:447), rewrote {rewrite.hits}")
    ns = {
        "torch": torch, "nn": nn, "F": F, "IntermediateLayerGetter": IntermediateLayerGetter,
        "clip": clip_pkg.clip, "_tokenizer": clip_pkg.simple_tokenizer.SimpleTokenizer(),
        "caption_text_feats": caption_text_feats, "__name__": "_lecb_ref_trainer",
    }
    exec(code, ns)
    return ns


def adapter_trainer_classes():
    """-> namespace with the reference `TextEncoder`, `AdapterTextEncoder`, `PromptLearner`, `Adapter`, `AdapterDenseCLIP` of
    trainers/Caption_distill_double_adapter.py (TA:41-457), AST-extracted and executed unmodified like `trainer_classes`."""
    import torch
    import torch.nn as nn
    from torch.nn import functional as F
    from torchvision.models._utils import IntermediateLayerGetter

    clip_pkg = clip_package()
    code = _extract(os.path.join(MC, "trainers", "Caption_distill_double_adapter.py"),
                    ["TextEncoder", "AdapterTextEncoder", "PromptLearner", "Adapter", "AdapterDenseCLIP"])
    ns = {"torch": torch, "nn": nn, "F": F, "IntermediateLayerGetter": IntermediateLayerGetter, "clip": clip_pkg.clip,
          "_tokenizer": clip_pkg.simple_tokenizer.SimpleTokenizer(), "__name__": "_lecb_ref_adapter_trainer"}
    exec(code, ns)
    return ns


def loss_functions():
    """-> namespace with the reference `ranking_loss`, `AsymmetricLoss_partial`, `dualcoop_loss`,
    `ASL_loss`, `ranking_loss_with_cooccurrence` (trainers/utils.py:85-190)."""
    import torch
    import torch.nn as nn
    from torch.nn import functional as F
    code = _extract(os.path.join(MC, "trainers", "utils.py"),
                    ["ranking_loss", "ranking_loss_with_cooccurrence", "AsymmetricLoss_partial",
                     "dualcoop_loss", "ASL_loss"])
    ns = {"torch": torch, "nn": nn, "F": F, "__name__": "_lecb_ref_losses"}
    exec(code, ns)
    return ns


def resample_loss_class():
    """-> the reference `ResampleLoss` (trainers/dbl.py:263-445) with the helpers it calls (dbl.py:20-65, 179-222).
    The module itself cannot be imported (mmcv, matplotlib, sklearn at dbl.py:4-9), and the class moves its frequency
    tables to the GPU in __init__ (`.cuda()`, dbl.py:326-341): the namespace gets an `mmcv.load` that unpickles and, for the
    duration of a construction, callers patch `torch.Tensor.cuda` to the identity (see oracle/make_golden.golden_resample)."""
    import functools
    import pickle
    import types
    import numpy as np
    import torch
    import torch.nn as nn
    from torch.nn import functional as F

    def _load(path):
        with open(path, "rb") as f:
            return pickle.load(f)

    code = _extract(os.path.join(MC, "trainers", "dbl.py"),
                    ["cross_entropy", "_expand_binary_labels", "binary_cross_entropy", "partial_cross_entropy", "reduce_loss",
                     "weight_reduce_loss", "ResampleLoss"])
    ns = {"torch": torch, "nn": nn, "F": F, "np": np, "functools": functools, "mmcv": types.SimpleNamespace(load=_load),
          "__name__": "_lecb_ref_dbl"}
    exec(code, ns)
    return ns["ResampleLoss"]


def checkpoint_functions():
    """-> namespace with the reference `save_checkpoint`, `load_checkpoint`, `load_pretrained_weights`
    (Dassl.pytorch-master/dassl/utils/torchtools.py:27-120, 266-320).  torch >= 2.6 made weights_only=True the default
    of torch.load, which cannot read the optimizer state these files hold: the namespace's `torch.load` passes False."""
    import pickle
    import shutil
    import types
    import warnings
    from collections import OrderedDict
    from functools import partial
    import torch
    path = os.path.join(MC, "Dassl.pytorch-master", "dassl", "utils", "torchtools.py")
    code = _extract(path, ["save_checkpoint", "load_checkpoint", "load_pretrained_weights"])
    tshim = types.SimpleNamespace(save=torch.save, cuda=torch.cuda,
                                  load=lambda *a, **k: torch.load(*a, **{"weights_only": False, **k}))
    ns = {"torch": tshim, "osp": os.path, "pickle": pickle, "shutil": shutil, "warnings": warnings, "partial": partial,
          "OrderedDict": OrderedDict, "mkdir_if_missing": lambda d: os.makedirs(d, exist_ok=True),
          "__name__": "_lecb_ref_torchtools"}
    exec(code, ns)
    return ns


def window_transform():
    """-> the reference `DatasetWrapperWithBlock._transform_image` (dassl/data/data_manager.py:348-492) as a plain function
    `f(self, tfm, img0)`; it only needs `self.k_tfm`, `self.multi_scale`, torch and torchvision's functional transforms."""
    import torch
    import torchvision.transforms.functional as F
    path = os.path.join(MC, "Dassl.pytorch-master", "dassl", "data", "data_manager.py")
    with open(path, "r") as f:
        tree = ast.parse(f.read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "DatasetWrapperWithBlock")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "_transform_image")
    mod = ast.Module(body=[fn], type_ignores=[])
    ast.fix_missing_locations(mod)
    ns = {"torch": torch, "F": F, "__name__": "_lecb_ref_windows"}
    exec(compile(mod, path, "exec"), ns)
    return ns["_transform_image"]


def fusion_functions(sims_scores):
    """-> namespace with the reference `fuse` / `fuse6` of gen_final_ans.py:18-71 (they read the module global
    `sims_scores`, provided here) and `adjust_predictions`, the helper nested in Caption_distill_double.test (T:611-615)."""
    import torch
    code = _extract(os.path.join(MC, "gen_final_ans.py"), ["fuse", "fuse6"])
    ns = {"torch": torch, "sims_scores": sims_scores, "__name__": "_lecb_ref_fusion"}
    exec(code, ns)
    with open(os.path.join(MC, "trainers", "Caption_distill_double.py"), "r") as f:
        tree = ast.parse(f.read())
    nested = [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "adjust_predictions"]
    if not nested:
        raise RuntimeError("adjust_predictions not found in the reference trainer")
    mod = ast.Module(body=[nested[0]], type_ignores=[])
    ast.fix_missing_locations(mod)
    exec(compile(mod, "Caption_distill_double.py", "exec"), ns)
    return ns


def mAP_function():
    """The numpy `mAP` of dassl/evaluation/evaluator.py:137-175 (AST-extracted: pure numpy)."""
    import numpy as np
    code = _extract(os.path.join(MC, "Dassl.pytorch-master", "dassl", "evaluation", "evaluator.py"),
                    ["average_precision", "mAP"])
    ns = {"np": np, "__name__": "_lecb_ref_map"}
    exec(code, ns)
    return ns["mAP"]


def coco_classnames():
    """`coco_object_categories` (datasets/data_helpers.py:252): first synonym of each COCO class."""
    path = os.path.join(MC, "datasets", "data_helpers.py")
    with open(path, "r") as f:
        tree = ast.parse(f.read())
    for node in tree.body:
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", "") == "coco_classname_synonyms":
            syn = ast.literal_eval(node.value)
            return [s[0] for s in syn]
    raise RuntimeError("coco_classname_synonyms not found")


class Cfg(dict):
    """dict-with-attributes standing in for the yacs CfgNode the reference reads."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def make_cfg(res, n_ctx=16, csc=False, use_evidence=False, ema=False, learn_scale=False,
             learn_spatial=False, spatial_text=50.0, spatial_image=50.0):
    """The keys DenseCLIP / PromptLearner read (SURVEY §5 'Config / flags'; defaults of
    train_caption.py:74-142 + the shipped yamls: spatial scales 50, fixed logit scale)."""
    return Cfg(
        TRAINER=Cfg(Caption=Cfg(N_CTX=n_ctx, CTX_INIT="", CSC=csc, CLASS_TOKEN_POSITION="end",
                                use_evidence=use_evidence, PREC="fp32")),
        INPUT=Cfg(SIZE=(res, res)),
        TRAIN=Cfg(IF_LEARN_SCALE=learn_scale, IF_LEARN_spatial_SCALE=learn_spatial,
                  spatial_SCALE_text=spatial_text, spatial_SCALE_image=spatial_image,
                  ema=ema, momentum=0.995, LOSSFUNC="double_ranking"),
    )


def build_reference_clip(arch, state_dict):
    """Reference `CLIP(...)` (clip/model.py:279) in fp32 eval mode with the synthetic weights."""
    M = clip_model_module()
    model = M.CLIP(*arch.ctor_args())
    missing, unexpected = model.load_state_dict(state_dict, strict=False)
    # attn_mask is a plain attribute, not a buffer; nothing else may be missing
    if missing or unexpected:
        raise RuntimeError(f"state_dict mismatch: missing={missing} unexpected={unexpected}")
    return model.float().eval()
