"""Model classes of the distillation path (reference export list: llavamod/model/__init__.py:17-23).
Other LLM families of the reference (llama, mistral, phi, minicpm, stablelm, gemma2, qwen-1.0, mpt) are out of scope."""
from .language_model.llava_qwen1_5 import LlavaQwen1_5ForCausalLM, LlavaQwen1_5Config  # noqa: F401
from .language_model.llava_qwen1_5_moe import (LLaVAMoDQwen1_5ForCausalLM, LLaVAMoDQwen1_5Config,  # noqa: F401
                                                LLaVAMoDQwen1_5ForCausalLMFineTune, EvalLLaVAMoDQwen1_5ForCausalLM)
from .language_model.llava_qwen2 import LlavaQwen2ForCausalLM, LlavaQwen2Config  # noqa: F401
from .language_model.llava_qwen2_moe import (LLaVAMoDQwen2ForCausalLM, LLaVAMoDQwen2Config,  # noqa: F401
                                              LLaVAMoDQwen2ForCausalLMFineTune, EvalLLaVAMoDQwen2ForCausalLM)
from .auto import AutoConfig, AutoModelForCausalLM  # noqa: F401,E402
