"""On-disk formats of the path's host side (SURVEY §8b "checkpoint format", §8f-3): prompt-learner checkpoints and the
test-time logit dumps, written and read exactly as the reference does so files move freely between the two code bases.

Reference: `dassl/utils/torchtools.py`:27-82 (`save_checkpoint`), :85-120 (`load_checkpoint`), :266-320
(`load_pretrained_weights`); `dassl/engine/trainer.py`:119-143 (`save_model`); trainers/Caption_distill_double.py:906-938
(`load_model`), :704-722 (`data*.pth` / `sim_matrix_B.pth` dumps read by gen_final_ans.py:14,74-78).  Pure host code: no kernels."""
from __future__ import annotations

import os
import os.path as osp
import shutil
import warnings
import weakref
from collections import OrderedDict

import torch

__all__ = ["register_owner", "save_checkpoint", "load_checkpoint", "load_pretrained_weights", "save_model", "load_model",
           "save_logit_dump", "load_logit_dump", "save_sim_matrix"]


def save_checkpoint(state, save_dir, is_best=False, remove_module_from_keys=True, model_name=""):
    """torchtools.py:27-82: `<save_dir>/<model_name or 'model.pth.tar-<epoch>'>`, a `checkpoint` file naming the latest
    one, and a `model-best.pth.tar` copy when `is_best`.  DDP's "module." prefix is stripped from the keys."""
    os.makedirs(save_dir, exist_ok=True)
    if remove_module_from_keys:
        new_state_dict = OrderedDict()
        for k, v in state["state_dict"].items():
            new_state_dict[k[7:] if k.startswith("module.") else k] = v
        state["state_dict"] = new_state_dict
    epoch = state["epoch"]
    if not model_name:
        model_name = "model.pth.tar-" + str(epoch)
    fpath = osp.join(save_dir, model_name)
    torch.save(state, fpath)
    with open(osp.join(save_dir, "checkpoint"), "w+") as f:
        f.write("{}\n".format(osp.basename(fpath)))
    if is_best:
        shutil.copy(fpath, osp.join(osp.dirname(fpath), "model-best.pth.tar"))
    return fpath


def load_checkpoint(fpath):
    """torchtools.py:85-120: ValueError for None, FileNotFoundError for a missing file, latin1 retry for python2 pickles.
    (torch >= 2.6 defaults to weights_only=True, which rejects the optimizer / scheduler state the reference stores in
    the same file, so the load is explicit about weights_only=False — these are the user's own training artefacts.)"""
    if fpath is None:
        raise ValueError("File path is None")
    if not osp.exists(fpath):
        raise FileNotFoundError('File is not found at "{}"'.format(fpath))
    map_location = None if torch.cuda.is_available() else "cpu"
    try:
        return torch.load(fpath, map_location=map_location, weights_only=False)
    except UnicodeDecodeError:          # python2-era pickle: the reference retries with latin1 (torchtools.py:108-114)
        return torch.load(fpath, map_location=map_location, weights_only=False, encoding="latin1")


def load_pretrained_weights(model, weight_path):
    """torchtools.py:266-320 (`cfg.MODEL.INIT_WEIGHTS`, T:767-768): layers unmatched in name or size are ignored,
    "module." prefixes are dropped.  Returns (matched, discarded) key lists."""
    checkpoint = load_checkpoint(weight_path)
    state_dict = checkpoint["state_dict"] if "state_dict" in checkpoint else checkpoint
    model_dict = model.state_dict()
    new_state_dict = OrderedDict()
    matched, discarded = [], []
    for k, v in state_dict.items():
        if k.startswith("module."):
            k = k[7:]
        if k in model_dict and model_dict[k].size() == v.size():
            new_state_dict[k] = v
            matched.append(k)
        else:
            discarded.append(k)
    model_dict.update(new_state_dict)
    model.load_state_dict(model_dict)
    if not matched:
        warnings.warn('The pretrained weights "{}" cannot be loaded, please check the key names manually '
                      "(** ignored and continue **)".format(weight_path))
    _invalidate(model)
    return matched, discarded


_OWNERS = weakref.WeakKeyDictionary()          # prompt learner -> weakref(the DenseCLIP module that caches its text features)


def register_owner(module, owner):
    """The trainer hands the prompt learner alone to save_model / load_model (T:774); loading into it must invalidate the
    text features its DenseCLIP module caches (T:421-439 never refreshes them)."""
    _OWNERS[module] = weakref.ref(owner)


def _invalidate(module):
    ref = _OWNERS.get(module)
    for m in (module, ref() if ref is not None else None):
        fn = getattr(m, "reset_prompt_cache", None)
        if callable(fn):
            fn()


def save_model(models, epoch, directory, is_best=False, model_name=""):
    """trainer.py:119-143.  `models`: {name: module} or {name: (module, optimizer, scheduler)} — the trainer registers
    each DenseCLIP's `prompt_learner` under its model name (T:774).  Stores `epoch + 1` like the reference."""
    paths = {}
    for name, entry in models.items():
        module, optim, sched = entry if isinstance(entry, (tuple, list)) else (entry, None, None)
        paths[name] = save_checkpoint(
            {"state_dict": module.state_dict(),
             "epoch": epoch + 1,
             "optimizer": None if optim is None else optim.state_dict(),
             "scheduler": None if sched is None else sched.state_dict()},
            osp.join(directory, name), is_best=is_best, model_name=model_name)
    return paths


def load_model(models, directory, epoch=None):
    """T:906-938: `<directory>/<name>/model.pth.tar[-epoch]`; the fixed token buffers are dropped from the file's state
    dict and the rest is loaded with strict=False.  Returns {name: stored epoch}; skipped (None) without a directory."""
    if not directory:
        print("Note that load_model() is skipped as no pretrained model is given")
        return None
    model_file = "model.pth.tar" if epoch is None else "model.pth.tar-" + str(epoch)
    epochs = {}
    for name, entry in models.items():
        module = entry[0] if isinstance(entry, (tuple, list)) else entry
        model_path = osp.join(directory, name, model_file)
        if not osp.exists(model_path):
            raise FileNotFoundError('Model not found at "{}"'.format(model_path))
        checkpoint = load_checkpoint(model_path)
        state_dict = checkpoint["state_dict"]
        for fixed in ("token_prefix", "token_suffix"):
            state_dict.pop(fixed, None)
        module.load_state_dict(state_dict, strict=False)
        _invalidate(module)
        epochs[name] = checkpoint["epoch"]
    return epochs


# ---------------------------------------------------------------- test-time dumps (T:704-722)
_DUMP_KEYS = ("output", "output_pos", "output_blocks", "output_pos_blocks")


def save_logit_dump(path, per_model):
    """`cfg.TEST.save_name` file: {model name: {"output": [N,K], "output_pos": [N,K] (+ "output_blocks": [N,nb,K],
    "output_pos_blocks": [N,nb,K] with sliding windows)}} of CPU tensors; lists of per-batch tensors are concatenated."""
    need_save = {}
    for name, d in per_model.items():
        need_save[name] = {}
        for k in _DUMP_KEYS:
            if k in d and d[k] is not None:
                v = d[k]
                v = torch.cat([t.detach().cpu() for t in v]) if isinstance(v, (list, tuple)) else v.detach().cpu()
                need_save[name][k] = v
        missing = [k for k in _DUMP_KEYS[:2] if k not in need_save[name]]
        if missing:
            raise KeyError(f"logit dump of model {name!r} lacks {missing}")
    os.makedirs(osp.dirname(osp.abspath(path)), exist_ok=True)
    torch.save(need_save, path)
    return need_save


def load_logit_dump(path):
    """gen_final_ans.py:74-78."""
    return torch.load(path, map_location="cpu", weights_only=False)


def save_sim_matrix(path, sims_all, sims_blocks_all, overwrite=False):
    """`sim_matrix_B.pth` (T:706-711, read at gen_final_ans.py:14): the retrieval top-k scores of the whole images
    [N,k] and of the windows [N,nb,k].  Like the reference, an existing file is kept unless `overwrite`."""
    cat = lambda v: torch.cat([t.detach().cpu() for t in v]) if isinstance(v, (list, tuple)) else v.detach().cpu()
    sim_matrix = {"sims_all": cat(sims_all), "sims_blocks_all": cat(sims_blocks_all)}
    if overwrite or not osp.exists(path):
        os.makedirs(osp.dirname(osp.abspath(path)), exist_ok=True)
        torch.save(sim_matrix, path)
    return sim_matrix
