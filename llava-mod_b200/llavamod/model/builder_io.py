"""Checkpoint IO in the reference's on-disk layout: ``config.json`` + ``pytorch_model.bin`` (or sharded / safetensors)
with the key names of SURVEY.md section 8b (reference: llavamod/train/align_train.py:623-631)."""
import glob
import json
import os

import torch


def load_state_dict_files(path):
    sd = {}
    idx = os.path.join(path, "pytorch_model.bin.index.json")
    files = []
    if os.path.exists(idx):
        with open(idx) as f:
            files = sorted(set(json.load(f)["weight_map"].values()))
    elif os.path.exists(os.path.join(path, "pytorch_model.bin")):
        files = ["pytorch_model.bin"]
    for fn in files:
        sd.update(torch.load(os.path.join(path, fn), map_location="cpu", weights_only=True))
    st = sorted(glob.glob(os.path.join(path, "*.safetensors")))
    if st and not sd:
        try:
            from safetensors.torch import load_file
        except Exception as e:  # pragma: no cover
            raise RuntimeError("safetensors checkpoint found but the safetensors package is missing") from e
        for fn in st:
            sd.update(load_file(fn))
    return sd


def load_into(module, sd, strict=True, prefix_strip=("base_model.", )):
    """Copies tensors into existing parameters IN PLACE (fused buffers stay fused)."""
    own = dict(module.named_parameters())
    own.update(dict(module.named_buffers()))
    missing, unexpected = [], []
    for k, v in sd.items():
        kk = k
        for p in prefix_strip:
            if kk.startswith(p):
                kk = kk[len(p):]
        if kk in own:
            with torch.no_grad():
                own[kk].copy_(v.to(own[kk].dtype))
        else:
            unexpected.append(k)
    for k in own:
        if k not in sd:
            missing.append(k)
    if strict and (missing or unexpected):
        raise KeyError("load_into: missing=%s unexpected=%s" % (missing[:8], unexpected[:8]))
    return missing, unexpected
