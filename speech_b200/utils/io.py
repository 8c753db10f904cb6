"""Checkpoints - host-side mirror of speech/utils/io.py:5-26 (reference), SURVEY.md section 8f
rank 3.

`save` / `load` keep the reference's on-disk contract exactly (`<path>/[tag_]model` = the whole
pickled module, `<path>/[tag_]preproc.pyc` = the pickled preprocessor; train.py:115-121,
eval.py:26), so checkpoints written by either side are read by the other as long as the class
path resolves.  `save_state` / `load_state` add what the reference lacks for resuming a run: the
state_dict (robust against class moves and torch's `weights_only` default), the optimiser state
and the epoch / iteration counters, in `<path>/[tag_]state`.
"""
import os
import pickle

import torch

MODEL = "model"
PREPROC = "preproc.pyc"
STATE = "state"


def get_names(path, tag):
    tag = tag + "_" if tag else ""
    return os.path.join(path, tag + MODEL), os.path.join(path, tag + PREPROC)


def save(model, preproc, path, tag=""):
    model_n, preproc_n = get_names(path, tag)
    torch.save(model, model_n)
    with open(preproc_n, "wb") as fid:
        pickle.dump(preproc, fid)


def load(path, tag=""):
    model_n, preproc_n = get_names(path, tag)
    model = torch.load(model_n, weights_only=False)     # whole-module pickle, as the reference
    with open(preproc_n, "rb") as fid:
        preproc = pickle.load(fid)
    return model, preproc


def save_state(model, path, tag="", optimizer=None, **counters):
    """state_dict (+ optimiser state and counters such as epoch=, iteration=) for resume."""
    blob = {"model": {k: v.detach().cpu() for k, v in model.state_dict().items()},
            "counters": dict(counters)}
    if optimizer is not None and hasattr(optimizer, "state_dict"):
        blob["optimizer"] = optimizer.state_dict()
    name = os.path.join(path, (tag + "_" if tag else "") + STATE)
    tmp = name + ".tmp"
    torch.save(blob, tmp)
    os.replace(tmp, name)            # never leave a half-written checkpoint behind
    return name


def load_state(model, path, tag="", optimizer=None):
    """restores `model` (and `optimizer`) in place; returns the saved counters."""
    name = os.path.join(path, (tag + "_" if tag else "") + STATE)
    blob = torch.load(name, map_location="cpu", weights_only=True)
    model.load_state_dict(blob["model"])
    if optimizer is not None and "optimizer" in blob and hasattr(optimizer, "load_state_dict"):
        optimizer.load_state_dict(blob["optimizer"])
    return blob.get("counters", {})
