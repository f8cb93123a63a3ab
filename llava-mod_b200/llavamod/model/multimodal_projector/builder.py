"""mm_projector: ``image_spatial_proj`` = Linear / mlp{N}x_gelu (reference: multimodal_projector/builder.py:26-66,125-149).
The qformer / simple / pool / video projectors are unused by the distillation shells and not built."""
import re

import torch
import torch.nn as nn

from ... import kernels as K
from ..language_model.qwen2_core import ParamLinear


class _Seq(nn.Module):
    """nn.Sequential-compatible key layout: Linear at even indices, GELU (parameter-free) at odd ones."""

    def __init__(self, linears):
        super().__init__()
        for j, lin in enumerate(linears):
            self.add_module(str(2 * j), lin)
        self.depth = len(linears)

    def linears(self):
        return [getattr(self, str(2 * j)) for j in range(self.depth)]


def build_image_projector(config, device="cuda", dtype=torch.bfloat16):
    projector_type = getattr(config, "image_projector_type", "linear") or "linear"
    std = getattr(config, "initializer_range", 0.02)

    def lin(i, o):
        w = torch.empty(o, i, device=device, dtype=dtype).normal_(0.0, std)
        return ParamLinear(w, torch.zeros(o, device=device, dtype=dtype))

    if projector_type == "linear":
        return _Seq([lin(config.mm_hidden_size, config.hidden_size)])
    m = re.match(r"^mlp(\d+)x_gelu$", projector_type)
    if m:
        depth = int(m.group(1))
        return _Seq([lin(config.mm_hidden_size, config.hidden_size)] + [lin(config.hidden_size, config.hidden_size) for _ in range(1, depth)])
    raise NotImplementedError(f"projector type {projector_type!r} is outside the distillation hot path (SURVEY.md 2.1 row 9)")


class build_projector(nn.Module):
    def __init__(self, config, delay_load=False, device="cuda", dtype=torch.bfloat16, **kwargs):
        super().__init__()
        self.image_spatial_proj = build_image_projector(config, device, dtype) if getattr(config, "mm_image_tower", None) is not None else None
        self.grad_views = {}

    def forward_image(self, x):
        """Linear (+bias) -> GELU -> Linear ... ; trainable, so every op carries a backward."""
        lins = self.image_spatial_proj.linears()
        for j, lin in enumerate(lins):
            if j > 0:
                x = K.gelu(x)
            x = K.linear(x, lin.weight, lin.bias, self.grad_views.get(id(lin.weight)), self.grad_views.get(id(lin.bias)))
        return x
