"""Sparse-MoE LLaVA-Qwen2 student.
Reference: llavamod/model/language_model/llava_qwen2_moe.py (LLaVAMoDQwen2Config :48-81, LLaVAMoDQwen2ForCausalLM :342-560,
...FineTune :564-626, Eval... :629-681)."""
from .llava_qwen_common import (LLaVAMoDFineTuneBase, LLaVAMoDForCausalLMBase, LlavaQwenModelBase, make_moe_config)

LLaVAMoDQwen2Config = make_moe_config("LLaVAMoDQwen2Config", "moe_llava_qwen2")


class LLaVAMoDQwen2Model(LlavaQwenModelBase):
    config_class = LLaVAMoDQwen2Config


class LLaVAMoDQwen2ForCausalLM(LLaVAMoDForCausalLMBase):
    config_class = LLaVAMoDQwen2Config
    model_class = LLaVAMoDQwen2Model


class LLaVAMoDQwen2ForCausalLMFineTune(LLaVAMoDFineTuneBase):
    config_class = LLaVAMoDQwen2Config
    model_class = LLaVAMoDQwen2Model


class EvalLLaVAMoDQwen2ForCausalLM(LLaVAMoDFineTuneBase):
    """Inference-time class: same construction from config.moe; routing uses eval_capacity_factor in eval()."""
    config_class = LLaVAMoDQwen2Config
    model_class = LLaVAMoDQwen2Model


# the reference's auto-factory registrations (llava_qwen2_moe.py:684-687), on this package's own registry
from ..auto import AutoConfig, AutoModelForCausalLM  # noqa: E402

AutoConfig.register("moe_llava_qwen2", LLaVAMoDQwen2Config)
AutoModelForCausalLM.register(LLaVAMoDQwen2Config, LLaVAMoDQwen2ForCausalLM)
AutoModelForCausalLM.register(LLaVAMoDQwen2Config, LLaVAMoDQwen2ForCausalLMFineTune)
AutoModelForCausalLM.register(LLaVAMoDQwen2Config, EvalLLaVAMoDQwen2ForCausalLM)
