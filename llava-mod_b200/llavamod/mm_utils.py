"""Image / prompt helpers on the data side of the hot path (reference: llavamod/mm_utils.py:13-62,65-71)."""
import torch

from .constants import IMAGE_TOKEN_INDEX


def expand2square(pil_img, background_color):
    """Pad to a centred square with the processor's mean colour (mm_utils.py:13-25; used when --image_aspect_ratio pad)."""
    from PIL import Image
    w, h = pil_img.size
    if w == h:
        return pil_img
    side = max(w, h)
    canvas = Image.new(pil_img.mode, (side, side), background_color)
    canvas.paste(pil_img, ((side - w) // 2, (side - h) // 2))
    return canvas


def process_images(images, image_processor, model_cfg):
    """mm_utils.py:28-40: per-image 'pad' path, else one batched processor call."""
    if getattr(model_cfg, "image_aspect_ratio", None) != "pad":
        return image_processor(images, return_tensors="pt")["pixel_values"]
    fill = tuple(int(c * 255) for c in image_processor.image_mean)
    out = [image_processor.preprocess(expand2square(im, fill), return_tensors="pt")["pixel_values"][0] for im in images]
    return torch.stack(out, dim=0) if all(o.shape == out[0].shape for o in out) else out


def tokenizer_image_token(prompt, tokenizer, image_token_index=IMAGE_TOKEN_INDEX, return_tensors=None):
    """Tokenise the text between '<image>' markers separately and put `image_token_index` (-200) where each marker stood
    (mm_utils.py:43-62).  If the tokenizer prepends BOS, it is kept once at the front and stripped from the later pieces."""
    pieces = [tokenizer(piece).input_ids for piece in prompt.split("<image>")]
    bos = getattr(tokenizer, "bos_token_id", None)
    has_bos = bool(pieces) and len(pieces[0]) > 0 and bos is not None and pieces[0][0] == bos
    ids = [pieces[0][0]] if has_bos else []
    skip = 1 if has_bos else 0
    for n, piece in enumerate(pieces):
        if n:
            ids.append(image_token_index)
        ids.extend(piece[skip:])
    if return_tensors is None:
        return ids
    if return_tensors == "pt":
        return torch.tensor(ids, dtype=torch.long)
    raise ValueError("Unsupported tensor type: %s" % return_tensors)


def get_model_name_from_path(model_path):
    parts = model_path.strip("/").split("/")
    return parts[-2] + "_" + parts[-1] if parts[-1].startswith("checkpoint-") else parts[-1]
