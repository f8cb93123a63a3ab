"""`AutoConfig` / `AutoModelForCausalLM` for the path's model families.

The reference registers its classes with transformers' auto factories (llava_qwen1_5.py:170-171, llava_qwen2.py:133-134,
llava_qwen1_5_moe.py:684-687, llava_qwen2_moe.py:684-687) so that `AutoConfig.from_pretrained(dir)` / `AutoModelForCausalLM.from_pretrained(dir)`
pick the class from `config.json`'s `model_type`.  The configs here are plain classes (no transformers import on the hot path), so the
registry is this module's own, with the same two entry points and the same `register` calls -- the LAST model class registered for a
config wins, as in transformers' mapping (for the MoE configs that is the Eval class, llava_qwen1_5_moe.py:687)."""
import json
import os

_CONFIGS = {}
_MODELS = {}


class AutoConfig:
    @staticmethod
    def register(model_type, config_class):
        if getattr(config_class, "model_type", model_type) != model_type:
            raise ValueError("config class %s has model_type %r, registered as %r" % (config_class.__name__, config_class.model_type, model_type))
        _CONFIGS[model_type] = config_class

    @staticmethod
    def for_model(model_type, **kw):
        if model_type not in _CONFIGS:
            raise ValueError("unknown model_type %r (registered: %s)" % (model_type, sorted(_CONFIGS)))
        return _CONFIGS[model_type](**kw)

    @staticmethod
    def from_pretrained(path, **kw):
        with open(os.path.join(path, "config.json")) as f:
            d = json.load(f)
        mt = d.get("model_type")
        if mt not in _CONFIGS:
            raise ValueError("%s/config.json has model_type %r; registered: %s" % (path, mt, sorted(_CONFIGS)))
        d.update(kw)
        return _CONFIGS[mt].from_dict(d)


class AutoModelForCausalLM:
    @staticmethod
    def register(config_class, model_class):
        _MODELS[config_class] = model_class

    @staticmethod
    def from_config(config, **kw):
        return _MODELS[type(config)](config, **kw)

    @staticmethod
    def from_pretrained(path, **kw):
        cfg = kw.pop("config", None) or AutoConfig.from_pretrained(path)
        if type(cfg) not in _MODELS:
            raise ValueError("no model class registered for %s" % type(cfg).__name__)
        return _MODELS[type(cfg)].from_pretrained(path, config=cfg, **kw)
