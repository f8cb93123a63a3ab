"""Dense LLaVA-Qwen2 wrapper (teacher, or dense student of the dense->dense stage).
Reference: llavamod/model/language_model/llava_qwen2.py (LlavaQwen2Config / LlavaQwen2Model / LlavaQwen2ForCausalLM)."""
from .llava_qwen_common import LlavaQwenForCausalLMBase, LlavaQwenModelBase
from .qwen2_core import Qwen2Config


class LlavaQwen2Config(Qwen2Config):
    model_type = "llava_qwen2"


class LlavaQwen2Model(LlavaQwenModelBase):
    config_class = LlavaQwen2Config


class LlavaQwen2ForCausalLM(LlavaQwenForCausalLMBase):
    config_class = LlavaQwen2Config
    model_class = LlavaQwen2Model


# the reference's auto-factory registrations (llava_qwen2.py:133-134), on this package's own registry
from ..auto import AutoConfig, AutoModelForCausalLM  # noqa: E402

AutoConfig.register("llava_qwen2", LlavaQwen2Config)
AutoModelForCausalLM.register(LlavaQwen2Config, LlavaQwen2ForCausalLM)
