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


# the reference's auto-factory registrations (llava_qwen1_5_moe.py:684-687), on this package's own registry
from ..auto import AutoConfig, AutoModelForCausalLM  # noqa: E402

AutoConfig.register("moe_llava_qwen1_5", LLaVAMoDQwen1_5Config)
AutoModelForCausalLM.register(LLaVAMoDQwen1_5Config, LLaVAMoDQwen1_5ForCausalLM)
AutoModelForCausalLM.register(LLaVAMoDQwen1_5Config, LLaVAMoDQwen1_5ForCausalLMFineTune)
AutoModelForCausalLM.register(LLaVAMoDQwen1_5Config, EvalLLaVAMoDQwen1_5ForCausalLM)
