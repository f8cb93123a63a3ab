"""Dense LLaVA-Qwen1_5 wrapper (teacher, or dense student of the dense->dense stage).
Reference: llavamod/model/language_model/llava_qwen1_5.py (LlavaQwen1_5Config / LlavaQwen1_5Model / LlavaQwen1_5ForCausalLM)."""
from .llava_qwen_common import LlavaQwenForCausalLMBase, LlavaQwenModelBase
from .qwen2_core import Qwen2Config


class LlavaQwen1_5Config(Qwen2Config):
    model_type = "llava_qwen1_5"


class LlavaQwen1_5Model(LlavaQwenModelBase):
    config_class = LlavaQwen1_5Config


class LlavaQwen1_5ForCausalLM(LlavaQwenForCausalLMBase):
    config_class = LlavaQwen1_5Config
    model_class = LlavaQwen1_5Model


# the reference's auto-factory registrations (llava_qwen1_5.py:170-171), on this package's own registry
from ..auto import AutoConfig, AutoModelForCausalLM  # noqa: E402

AutoConfig.register("llava_qwen1_5", LlavaQwen1_5Config)
AutoModelForCausalLM.register(LlavaQwen1_5Config, LlavaQwen1_5ForCausalLM)
