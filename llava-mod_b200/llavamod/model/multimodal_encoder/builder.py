"""Vision tower dispatch by substring of the tower name (reference: multimodal_encoder/builder.py:15-36)."""
from .clip_encoder import CLIPVisionTower, CLIPVisionConfig


def build_image_tower(image_tower_cfg, **kwargs):
    image_tower = getattr(image_tower_cfg, "mm_image_tower", getattr(image_tower_cfg, "image_tower", None))
    if isinstance(image_tower, (dict, CLIPVisionConfig)):
        return CLIPVisionTower(image_tower, args=image_tower_cfg, **kwargs)
    if getattr(image_tower_cfg, "s2", False):
        raise NotImplementedError("S2 multi-scale towers are outside the distillation hot path (SURVEY.md 2.1 row 8)")
    if "openai" in image_tower or "laion" in image_tower:
        return CLIPVisionTower(image_tower, args=image_tower_cfg, cache_dir="./cache_dir", **kwargs)
    if "google" in image_tower:
        raise NotImplementedError("SigLIP towers are outside the distillation hot path (SURVEY.md 2.1 row 8)")
    raise ValueError(f"Unknown image tower: {image_tower}")
