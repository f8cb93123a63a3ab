"""Output containers (reference: llavamod/model/utils.py:120-127, llava_qwen1_5_moe.py:91-109) and small helpers
(create_reference_model / disable_dropout_in_model, llavamod/model/utils.py:34-112)."""
from dataclasses import dataclass, fields
from typing import Any, Optional, Tuple

import torch


class _Output:
    """Attribute + index access like transformers.ModelOutput, without importing transformers."""

    def __getitem__(self, k):
        if isinstance(k, str):
            return getattr(self, k)
        vals = [getattr(self, f.name) for f in fields(self) if getattr(self, f.name) is not None]
        return vals[k]

    def keys(self):
        return [f.name for f in fields(self) if getattr(self, f.name) is not None]

    def __contains__(self, k):
        return k in self.keys()


@dataclass
class CausalLMOutputWithPast(_Output):
    loss: Optional[torch.Tensor] = None
    logits: Optional[torch.Tensor] = None
    labels: Optional[torch.Tensor] = None
    past_key_values: Optional[Any] = None
    hidden_states: Optional[Tuple[torch.Tensor]] = None
    attentions: Optional[Tuple[torch.Tensor]] = None


@dataclass
class MoECausalLMOutputWithPast(_Output):
    loss: Optional[torch.Tensor] = None
    moe_loss: Optional[torch.Tensor] = None
    logits: Optional[torch.Tensor] = None
    labels: Optional[torch.Tensor] = None
    past_key_values: Optional[Any] = None
    hidden_states: Optional[Tuple[torch.Tensor]] = None
    attentions: Optional[Tuple[torch.Tensor]] = None
    moe_loss_list: Optional[Tuple[torch.Tensor]] = None


def disable_dropout_in_model(model: torch.nn.Module) -> None:
    for module in model.modules():
        if isinstance(module, torch.nn.Dropout):
            module.p = 0


def create_reference_model(model):
    """Frozen deep copy in eval mode (reference: llavamod/model/utils.py:34-112, without the shared-layer option)."""
    import copy
    ref = copy.deepcopy(model)
    for p in ref.parameters():
        p.requires_grad = False
    return ref.eval()
