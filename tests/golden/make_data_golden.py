"""Generates tests/golden/data_pipeline.pt (+ the tiny tokenizer and two tiny images it uses) by running the REFERENCE's own data
pipeline in this container:  python tests/golden/make_data_golden.py

What runs is the reference's code, loaded in place from /root/reference (read-only): llavamod/data/data_utils.py (preprocess_multimodal,
preprocess_phi, preprocess_plain), llavamod/data/dataset.py (LazySupervisedDataset, LazyDPODataset and both collators),
llavamod/mm_utils.py (tokenizer_image_token, expand2square), llavamod/conversation.py, and the sampler functions of
llavamod/train/align_trainer.py:68-163 (that file itself cannot be imported here -- accelerate / transformers 4.37 -- so exactly those
lines are exec'd).  `llavamod.model` is stubbed: data_utils only takes the name `transformers` from its star import.

The tokenizer is a byte-level BPE (the family Qwen's tokenizer belongs to: no BOS, leading-space merges) trained offline on a few
sentences; it is committed as tests/golden/tiny_tokenizer.json so the tests tokenise exactly like the golden run did."""
import json
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LLAVAMOD_REFERENCE", "/root/reference")

CORPUS = [
    "A chat between a curious user and an artificial intelligence assistant.",
    "The assistant gives helpful, detailed, and polite answers to the user's questions.",
    "USER: What is shown in the image? ASSISTANT: A small red square on a green field.",
    "USER: Describe the picture briefly. ASSISTANT: Two birds sit on a wire above the street.",
    "USER: How many birds are there? ASSISTANT: There are two birds.",
    "Provide a brief description of the given image. a cartoon illustration of a winged buffalo with an angry expression .",
    "Is the buffalo angry? Yes, it looks angry. No, it looks calm and friendly.",
]


def build_tokenizer(path):
    from tokenizers import Tokenizer, decoders, models, pre_tokenizers, trainers
    tok = Tokenizer(models.BPE())
    tok.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False)
    tok.decoder = decoders.ByteLevel()
    trainer = trainers.BpeTrainer(vocab_size=420, special_tokens=["<|endoftext|>", "<|extra_0|>"],
                                  initial_alphabet=pre_tokenizers.ByteLevel.alphabet())
    tok.train_from_iterator(CORPUS, trainer)
    tok.save(path)


def load_tokenizer(path, model_max_length=128):
    from transformers import PreTrainedTokenizerFast
    t = PreTrainedTokenizerFast(tokenizer_file=path, eos_token="<|endoftext|>", unk_token="<|extra_0|>", model_max_length=model_max_length,
                                padding_side="right")
    t.pad_token = t.unk_token          # align_train.py:436
    return t


def make_images(folder):
    from PIL import Image
    os.makedirs(folder, exist_ok=True)
    Image.new("RGB", (40, 24), (200, 30, 30)).save(os.path.join(folder, "wide.png"))
    im = Image.new("RGB", (20, 36), (20, 160, 60))
    im.putpixel((3, 5), (255, 255, 255))
    im.save(os.path.join(folder, "tall.png"))


SFT_RECORDS = [
    {"image": "wide.png", "conversations": [
        {"from": "human", "value": "<image>\nWhat is shown in the image?"},
        {"from": "gpt", "value": "A small red square on a green field."},
        {"from": "human", "value": "How many birds are there?"},
        {"from": "gpt", "value": "There are two birds."}]},
    {"image": ["tall.png", "wide.png"], "conversations": [
        {"from": "human", "value": "<image><image>\nDescribe the picture briefly."},
        {"from": "gpt", "value": "Two birds sit on a wire above the street."}]},
    {"conversations": [
        {"from": "human", "value": "Is the buffalo angry?"},
        {"from": "gpt", "value": "Yes, it looks angry."}]},
    {"image": "missing_file.png", "conversations": [
        {"from": "gpt", "value": "(a leading assistant turn is dropped)"},
        {"from": "human", "value": "<image>\nDescribe the picture briefly."},
        {"from": "gpt", "value": "A small red square."}]},
]
DPO_RECORDS = [
    {"image": "tall.png",
     "conversations": [{"from": "human", "value": "<image>\nIs the buffalo angry?"}, {"from": "gpt", "value": "Yes, it looks angry."}],
     "chosen": [{"from": "human", "value": "<image>\nIs the buffalo angry?"}, {"from": "gpt", "value": "Yes, it looks angry."}],
     "rejected": [{"from": "human", "value": "<image>\nIs the buffalo angry?"}, {"from": "gpt", "value": "No, it looks calm and friendly."}]},
    {"conversations": [{"from": "human", "value": "How many birds are there?"}, {"from": "gpt", "value": "There are two birds."}],
     "chosen": [{"from": "human", "value": "How many birds are there?"}, {"from": "gpt", "value": "There are two birds."}],
     "rejected": [{"from": "human", "value": "How many birds are there?"}, {"from": "gpt", "value": "A small red square."}]},
]
PLAIN_SOURCES = [[{"from": "human", "value": "Provide a brief description of the given image.\n<image>"},
                  {"from": "gpt", "value": "a cartoon illustration of a winged buffalo with an angry expression ."}]]
SAMPLER_CASES = [
    dict(lengths=[5, 9, 3, 12, 7, 8, 2, 11, 6, 4, 10, 1], batch_size=2, world_size=2, seed=0, modality=False),
    dict(lengths=[5, 9, 3, 12, 7, 8, 2, 11, 6, 4, 10], batch_size=2, world_size=2, seed=1, modality=False),
    dict(lengths=[5, -9, 3, 12, -7, 8, -2, 11, 6, -4, 10, 1, -13, 14, 15, -16, 17], batch_size=2, world_size=2, seed=2, modality=True),
]


def load_reference():
    import transformers
    base = os.path.join(REF, "llavamod")
    for name, path in (("llavamod", base), ("llavamod.data", os.path.join(base, "data"))):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
    stub = types.ModuleType("llavamod.model")
    stub.transformers = transformers
    sys.modules["llavamod.model"] = stub
    import llavamod.conversation as conv
    import llavamod.data.data_utils as du
    import llavamod.data.dataset as ds
    import llavamod.mm_utils as mu
    src = open(os.path.join(base, "train", "align_trainer.py")).read().split("\n")
    ns = {}
    exec("import torch\nfrom typing import List, Optional\nfrom torch.utils.data import Sampler\n" + "\n".join(src[67:163]), ns)
    return conv, du, ds, mu, ns


def main():
    tok_path = os.path.join(HERE, "tiny_tokenizer.json")
    img_dir = os.path.join(HERE, "data_imgs")
    if not os.path.exists(tok_path):
        build_tokenizer(tok_path)
    make_images(img_dir)
    with open(os.path.join(HERE, "data_sft.json"), "w") as f:
        json.dump(SFT_RECORDS, f, indent=1)
    with open(os.path.join(HERE, "data_dpo.json"), "w") as f:
        json.dump(DPO_RECORDS, f, indent=1)
    conv, du, ds, mu, sampler_ns = load_reference()
    from transformers import CLIPImageProcessor
    tok = load_tokenizer(tok_path)
    proc = CLIPImageProcessor(size={"shortest_edge": 32}, crop_size={"height": 32, "width": 32})
    out = {"meta": dict(pad_token_id=tok.pad_token_id, eos_token_id=tok.eos_token_id, vocab=len(tok))}
    ds.local_rank = 1          # silence rank0_print
    for aspect in ("square", "pad"):
        conv.default_conversation = conv.conv_templates["qwen"]
        args = types.SimpleNamespace(image_folder=img_dir, image_processor=proc, image_aspect_ratio=aspect, is_multimodal=True,
                                     mm_use_im_start_end=False, num_frames=8, data_path=[os.path.join(HERE, "data_sft.json")])
        sft = ds.LazySupervisedDataset(data_path=args.data_path, tokenizer=tok, data_args=args)
        items = [sft[i] for i in range(len(sft))]
        coll = ds.DataCollatorForSupervisedDataset(tokenizer=tok)
        batch = coll(items)
        out["sft_" + aspect] = dict(items=[{k: v for k, v in it.items()} for it in items], batch=batch, modality_lengths=sft.modality_lengths)
        args.data_path = [os.path.join(HERE, "data_dpo.json")]
        dpo = ds.LazyDPODataset(data_path=args.data_path, tokenizer=tok, data_args=args)
        ditems = [dpo[i] for i in range(len(dpo))]
        out["dpo_" + aspect] = dict(items=ditems, batch=ds.DataCollatorForDPODataset(tokenizer=tok)(ditems), modality_lengths=dpo.modality_lengths)
    # <im_start>/<im_end> wrapping and the plain (adaptor pre-training) template
    import copy
    args.mm_use_im_start_end = True
    out["mm_wrapped"] = du.preprocess_multimodal(copy.deepcopy([SFT_RECORDS[1]["conversations"]]), args)
    conv.default_conversation = conv.conv_templates["plain"]
    plain = du.preprocess(copy.deepcopy(PLAIN_SOURCES), tok, has_image=True)
    out["plain"] = dict(input_ids=[t.clone() for t in plain["input_ids"]], labels=[t.clone() for t in plain["labels"]])
    conv.default_conversation = conv.conv_templates["qwen"]
    out["tokenizer_image_token"] = {p: mu.tokenizer_image_token(p, tok) for p in ("<image>\nWhat is shown?", "no image here", "a<image>b<image>")}
    # sampler
    cases = []
    for c in SAMPLER_CASES:
        g = torch.Generator().manual_seed(c["seed"])
        torch.manual_seed(100 + c["seed"])      # the modality path draws its inner shuffles from the global RNG
        fn = sampler_ns["get_modality_length_grouped_indices" if c["modality"] else "get_length_grouped_indices"]
        cases.append(dict(c, indices=fn(c["lengths"], c["batch_size"], c["world_size"], generator=g)))
    out["sampler"] = cases
    torch.save(out, os.path.join(HERE, "data_pipeline.pt"))
    print("wrote", os.path.join(HERE, "data_pipeline.pt"), {k: type(v).__name__ for k, v in out.items()})


if __name__ == "__main__":
    main()
