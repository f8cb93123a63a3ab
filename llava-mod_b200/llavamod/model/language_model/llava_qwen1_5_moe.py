"""Sparse-MoE LLaVA-Qwen1_5 student.
Reference: llavamod/model/language_model/llava_qwen1_5_moe.py (LLaVAMoDQwen1_5Config :48-81, LLaVAMoDQwen1_5ForCausalLM :342-560,
...FineTune :564-626, Eval... :629-681)."""
from .llava_qwen_common import (LLaVAMoDFineTuneBase, LLaVAMoDForCausalLMBase, LlavaQwenModelBase, make_moe_config)

LLaVAMoDQwen1_5Config = make_moe_config("LLaVAMoDQwen1_5Config", "moe_llava_qwen1_5")


class LLaVAMoDQwen1_5Model(LlavaQwenModelBase):
    config_class = LLaVAMoDQwen1_5Config


class LLaVAMoDQwen1_5ForCausalLM(LLaVAMoDForCausalLMBase):
    config_class = LLaVAMoDQwen1_5Config
    model_class = LLaVAMoDQwen1_5Model


class LLaVAMoDQwen1_5ForCausalLMFineTune(LLaVAMoDFineTuneBase):
    config_class = LLaVAMoDQwen1_5Config
    model_class = LLaVAMoDQwen1_5Model


class EvalLLaVAMoDQwen1_5ForCausalLM(LLaVAMoDFineTuneBase):
    """Inference-time class: same construction from config.moe; routing uses eval_capacity_factor in eval()."""
    config_class = LLaVAMoDQwen1_5Config
    model_class = LLaVAMoDQwen1_5Model
