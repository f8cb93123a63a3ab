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


class KeywordsStoppingCriteria:
    """Stop when the newly generated tail decodes to (or ends with the ids of) one of `keywords` (mm_utils.py:73-105; a plain callable
    `(output_ids, scores) -> bool`, which is all `generate` needs -- no dependency on transformers.StoppingCriteria)."""

    def __init__(self, keywords, tokenizer, input_ids):
        self.keywords = keywords
        self.keyword_ids = []
        self.max_keyword_len = 0
        for kw in keywords:
            ids = tokenizer(kw).input_ids
            if len(ids) > 1 and ids[0] == getattr(tokenizer, "bos_token_id", None):
                ids = ids[1:]
            self.max_keyword_len = max(self.max_keyword_len, len(ids))
            self.keyword_ids.append(torch.tensor(ids))
        self.tokenizer = tokenizer
        self.start_len = input_ids.shape[1]

    def call_for_batch(self, output_ids, scores=None, **kw):
        offset = min(output_ids.shape[1] - self.start_len, self.max_keyword_len)
        for kid in self.keyword_ids:
            kid = kid.to(output_ids.device)
            if kid.numel() and output_ids.shape[1] >= kid.shape[0] and bool((output_ids[0, -kid.shape[0]:] == kid).all()):
                return True
        if offset <= 0:
            return False
        text = self.tokenizer.batch_decode(output_ids[:, -offset:], skip_special_tokens=True)[0]
        return any(kw_ in text for kw_ in self.keywords)

    def __call__(self, output_ids, scores=None, **kw):
        return all(self.call_for_batch(output_ids[i].unsqueeze(0), scores) for i in range(output_ids.shape[0]))
