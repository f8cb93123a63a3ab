"""Conversation -> (input_ids, labels) for the distillation recipes.

Reference: llavamod/data/data_utils.py -- preprocess_multimodal :102-151, preprocess_phi :318-394 (what `--version qwen` runs),
preprocess_plain :627-650, preprocess (dispatch) :653-711.  The label masks produced here ARE the KD / CE masks of the CUDA loss
kernels (lmod_kl_fwd_bwd reads `labels != -100`), so the token arithmetic below follows the reference to the token -- including its
quirks (the "+1 for eos" round length, the "-1" on the instruction length, all-ignored labels on a length mismatch)."""
import copy
from typing import Dict, Sequence

import torch

from .. import conversation as conversation_lib
from ..constants import (DEFAULT_IM_END_TOKEN, DEFAULT_IM_START_TOKEN, DEFAULT_IMAGE_TOKEN, DEFAULT_VID_END_TOKEN, DEFAULT_VID_START_TOKEN,
                         DEFAULT_VIDEO_TOKEN, IGNORE_INDEX, MAX_IMAGE_LENGTH, MAX_VIDEO_LENGTH)
from ..mm_utils import expand2square, tokenizer_image_token  # noqa: F401  (expand2square re-exported like the reference module)

_PHI_FAMILY = ("phi", "qwen", "stablelm")


def preprocess_multimodal(sources: Sequence, data_args) -> Sequence:
    """Normalise the media placeholders of every turn in place (data_utils.py:102-151): cap a leading run of <image> at
    MAX_IMAGE_LENGTH, refuse more than MAX_VIDEO_LENGTH <video>, optionally wrap with <im_start>/<im_end>, and expand each <video>
    into `num_frames` <image> tokens."""
    if not data_args.is_multimodal:
        return sources
    image_tok = DEFAULT_IMAGE_TOKEN
    video_tok = DEFAULT_IMAGE_TOKEN * data_args.num_frames
    if data_args.mm_use_im_start_end:
        image_tok = DEFAULT_IM_START_TOKEN + image_tok + DEFAULT_IM_END_TOKEN
        video_tok = DEFAULT_VID_START_TOKEN + video_tok + DEFAULT_VID_END_TOKEN
    for turns in sources:
        for turn in turns:
            text = turn["value"]
            if text.startswith(DEFAULT_IMAGE_TOKEN) or text.startswith(DEFAULT_VIDEO_TOKEN):
                if "mmtag" in conversation_lib.default_conversation.version:
                    text = text.replace(DEFAULT_IMAGE_TOKEN, "<Image>" + DEFAULT_IMAGE_TOKEN + "</Image>")
                n_img = text.count(DEFAULT_IMAGE_TOKEN)
                if n_img > MAX_IMAGE_LENGTH:
                    text = text.replace(DEFAULT_IMAGE_TOKEN * n_img, DEFAULT_IMAGE_TOKEN * MAX_IMAGE_LENGTH).strip()
                if text.count(DEFAULT_VIDEO_TOKEN) > MAX_VIDEO_LENGTH:
                    raise ValueError(text)
            turn["value"] = text.replace(DEFAULT_IMAGE_TOKEN, image_tok).replace(DEFAULT_VIDEO_TOKEN, video_tok)
    return sources


def _render_two_sep(sources, conv):
    """One prompt string per conversation; a leading non-human turn is dropped and the roles must alternate (data_utils.py:326-338)."""
    role_of = {"human": conv.roles[0], "gpt": conv.roles[1]}
    prompts = []
    for n, turns in enumerate(sources):
        if role_of[turns[0]["from"]] != conv.roles[0]:
            turns = turns[1:]
        conv.messages = []
        for k, turn in enumerate(turns):
            role = role_of[turn["from"]]
            assert role == conv.roles[k % 2], f"{n}"
            conv.append_message(role, turn["value"])
        prompts.append(conv.get_prompt())
    return prompts


def preprocess_phi(sources, tokenizer, has_image: bool = False) -> Dict:
    """Two-separator chat (`--version qwen | phi | stablelm`): labels keep only the assistant answers (+ the sep2 that ends them).

    Per round r = "... USER: q ASSISTANT: a" (split at sep2): round_len = tokens(r) + 1 (the sep2/eos token) and the first
    tokens(r up to and including "ASSISTANT: ") - 1 labels of the round are ignored; everything after the last round is ignored; if
    the accumulated length disagrees with the number of non-pad tokens, the whole sample is ignored with a warning
    (data_utils.py:353-390)."""
    conv = conversation_lib.default_conversation.copy()
    assert conv.sep_style == conversation_lib.SeparatorStyle.TWO
    prompts = _render_two_sep(sources, conv)

    def count(text):
        return len(tokenizer_image_token(text, tokenizer)) if has_image else len(tokenizer(text).input_ids)

    if has_image:
        input_ids = torch.stack([tokenizer_image_token(p, tokenizer, return_tensors="pt") for p in prompts], dim=0)
    else:
        input_ids = tokenizer(prompts, return_tensors="pt", padding="longest", max_length=tokenizer.model_max_length, truncation=True).input_ids
    labels = input_ids.clone()
    answer_mark = conv.sep + conv.roles[1] + ": "
    for prompt, row in zip(prompts, labels):
        n_tokens = int(row.ne(tokenizer.pad_token_id).sum())
        pos = 0
        for rnd in prompt.split(conv.sep2):
            if rnd == "":
                break
            halves = rnd.split(answer_mark)
            if len(halves) != 2:
                break
            row[pos: pos + count(halves[0] + answer_mark) - 1] = IGNORE_INDEX
            pos += count(rnd) + 1
        row[pos:] = IGNORE_INDEX
        if pos < tokenizer.model_max_length and pos != n_tokens:
            row[:] = IGNORE_INDEX
            print(f"WARNING: tokenization mismatch: {pos} vs. {n_tokens}. (ignored)")
    return dict(input_ids=input_ids, labels=labels)


def preprocess_plain(sources, tokenizer) -> Dict:
    """Adaptor pre-training (`--version plain`): the human turn collapses to a bare <image>, the caption (+ sep) is the target
    (data_utils.py:627-650)."""
    ids, labels = [], []
    for turns in sources:
        assert len(turns) == 2
        assert DEFAULT_IMAGE_TOKEN in turns[0]["value"]
        turns[0]["value"] = DEFAULT_IMAGE_TOKEN
        row = tokenizer_image_token(turns[0]["value"] + turns[1]["value"] + conversation_lib.default_conversation.sep, tokenizer, return_tensors="pt")
        tgt = row.clone()
        tgt[: len(tokenizer_image_token(turns[0]["value"], tokenizer))] = IGNORE_INDEX
        ids.append(row)
        labels.append(tgt)
    return dict(input_ids=ids, labels=labels)


def preprocess(sources, tokenizer, has_image: bool = False) -> Dict:
    """Dispatch on the active template (data_utils.py:653-677).  Only the templates of the Qwen recipes are carried."""
    conv = conversation_lib.default_conversation
    if conv.sep_style == conversation_lib.SeparatorStyle.PLAIN:
        return preprocess_plain(sources, tokenizer)
    if conv.version.startswith(_PHI_FAMILY):
        return preprocess_phi(sources, tokenizer, has_image=has_image)
    raise NotImplementedError("label masking for conversation version %r is outside the Qwen distillation path" % conv.version)


def deep_copy_turns(samples, key):
    return copy.deepcopy([s[key] for s in samples])
