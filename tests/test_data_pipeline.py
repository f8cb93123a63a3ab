"""SURVEY section 8f row N1 -- the data side of the path (llavamod/data, conversation, mm_utils, train/sampler) against golden
outputs of the REFERENCE's own pipeline run in the build container (tests/golden/make_data_golden.py -> data_pipeline.pt).
Integer work: token ids, label masks, attention masks and sampler orders must match bit for bit; pixel tensors exactly (same
CLIPImageProcessor, same expand2square)."""
import copy
import json
import os
import types

import pytest
import torch

from tests.golden.make_data_golden import PLAIN_SOURCES, SFT_RECORDS, load_tokenizer


@pytest.fixture(scope="module")
def env(golden_dir):
    from transformers import CLIPImageProcessor
    from llavamod import conversation as conversation_lib
    from llavamod.data import dataset as D
    D.local_rank = 1
    conversation_lib.set_default_conversation("qwen")
    g = torch.load(os.path.join(golden_dir, "data_pipeline.pt"), weights_only=False)
    tok = load_tokenizer(os.path.join(golden_dir, "tiny_tokenizer.json"))
    proc = CLIPImageProcessor(size={"shortest_edge": 32}, crop_size={"height": 32, "width": 32})

    def args(aspect, path):
        return types.SimpleNamespace(image_folder=os.path.join(golden_dir, "data_imgs"), image_processor=proc, image_aspect_ratio=aspect,
                                     is_multimodal=True, mm_use_im_start_end=False, num_frames=8, data_path=[os.path.join(golden_dir, path)])
    return types.SimpleNamespace(g=g, tok=tok, args=args, D=D)


def _same(a, b, path=""):
    if isinstance(b, torch.Tensor):
        assert isinstance(a, torch.Tensor) and a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b), path
    elif isinstance(b, dict):
        assert list(a.keys()) == list(b.keys()), (path, list(a.keys()), list(b.keys()))
        for k in b:
            _same(a[k], b[k], path + "/" + str(k))
    elif isinstance(b, (list, tuple)):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, path + "[%d]" % i)
    else:
        assert a == b, (path, a, b)


def test_tokenizer_has_no_bos_and_meta_matches(env):
    assert env.g["meta"] == dict(pad_token_id=env.tok.pad_token_id, eos_token_id=env.tok.eos_token_id, vocab=len(env.tok))


@pytest.mark.parametrize("aspect", ["square", "pad"])
def test_supervised_dataset_and_collator_match_reference(env, aspect, capsys):
    ds = env.D.LazySupervisedDataset(data_path=env.args(aspect, "data_sft.json").data_path, tokenizer=env.tok,
                                     data_args=env.args(aspect, "data_sft.json"))
    items = [ds[i] for i in range(len(ds))]
    ref = env.g["sft_" + aspect]
    _same(items, ref["items"], "items")
    assert ds.modality_lengths == ref["modality_lengths"]
    batch = env.D.DataCollatorForSupervisedDataset(tokenizer=env.tok)(items)
    _same(batch, ref["batch"], "batch")
    assert batch["attention_mask"].dtype == torch.bool and len(batch["images"]) == 5          # 1 + 2 + blank + fallback
    assert "using fallback image" in capsys.readouterr().out                                  # the unreadable file of record 3
    # the masks that reach the loss kernels: only assistant answers (and their closing <|endoftext|>) are supervised
    lab = items[0]["labels"]
    assert (lab != -100).sum() > 0 and lab[0] == -100 and lab[-1] == env.tok.eos_token_id


@pytest.mark.parametrize("aspect", ["square", "pad"])
def test_preference_dataset_and_collator_match_reference(env, aspect):
    a = env.args(aspect, "data_dpo.json")
    ds = env.D.LazyDPODataset(data_path=a.data_path, tokenizer=env.tok, data_args=a)
    items = [ds[i] for i in range(len(ds))]
    ref = env.g["dpo_" + aspect]
    _same(items, ref["items"], "items")
    _same(env.D.DataCollatorForDPODataset(tokenizer=env.tok)(items), ref["batch"], "batch")
    assert ds.modality_lengths == ref["modality_lengths"]


def test_im_start_end_wrapping_and_plain_template(env):
    from llavamod import conversation as conversation_lib
    from llavamod.data import data_utils as U
    a = env.args("square", "data_sft.json")
    a.mm_use_im_start_end = True
    assert U.preprocess_multimodal(copy.deepcopy([SFT_RECORDS[1]["conversations"]]), a) == env.g["mm_wrapped"]
    conversation_lib.set_default_conversation("plain")
    try:
        _same(U.preprocess(copy.deepcopy(PLAIN_SOURCES), env.tok, has_image=True), env.g["plain"], "plain")
    finally:
        conversation_lib.set_default_conversation("qwen")
    with pytest.raises(NotImplementedError):
        conversation_lib.set_default_conversation("llama_2")


def test_tokenizer_image_token(env):
    from llavamod.mm_utils import tokenizer_image_token
    for prompt, ids in env.g["tokenizer_image_token"].items():
        assert tokenizer_image_token(prompt, env.tok) == ids
        assert tokenizer_image_token(prompt, env.tok, return_tensors="pt").tolist() == ids
    with pytest.raises(ValueError):
        tokenizer_image_token("x", env.tok, return_tensors="np")


def test_length_mismatch_masks_the_whole_sample(env, capsys):
    """data_utils.py:384-390: when the per-round token arithmetic does not add up (here: the answer contains the round separator, so
    the split yields a piece that is not a full round), every label of the sample is ignored and a warning is printed."""
    from llavamod.data import data_utils as U
    src = [[{"from": "human", "value": "How many birds are there?"}, {"from": "gpt", "value": "There are<|endoftext|>two birds."}]]
    out = U.preprocess(src, env.tok, has_image=False)
    assert (out["labels"] == -100).all()
    assert "tokenization mismatch" in capsys.readouterr().out


def test_video_records_are_refused_and_resampled(env, tmp_path, capsys):
    recs = [{"video": "a.mp4", "conversations": SFT_RECORDS[2]["conversations"]}, SFT_RECORDS[2]]
    p = tmp_path / "v.json"
    p.write_text(json.dumps(recs))
    a = env.args("square", "data_sft.json")
    ds = env.D.LazySupervisedDataset(data_path=[str(p)], tokenizer=env.tok, data_args=a)
    torch.manual_seed(0)
    item = ds[0]                      # the video record fails -> a random other record is served, like the reference's except branch
    assert "LanguageBind" in capsys.readouterr().out and item["input_ids"].numel() > 0


def test_length_grouped_sampler_matches_reference(env):
    from llavamod.train import sampler as S
    for c in env.g["sampler"]:
        g = torch.Generator().manual_seed(c["seed"])
        torch.manual_seed(100 + c["seed"])
        fn = S.get_modality_length_grouped_indices if c["modality"] else S.get_length_grouped_indices
        assert fn(c["lengths"], c["batch_size"], c["world_size"], generator=g) == c["indices"]
    s = S.LengthGroupedSampler(2, 2, lengths=[3, 1, 2, 5, 4, 6, 8, 7], group_by_modality=True)
    order = list(s)
    assert sorted(order) == list(range(8)) and len(s) == 8
    with pytest.raises(ValueError):
        S.LengthGroupedSampler(2, 2)
    # per-rank shards: disjoint, equal length, whole per-device batches taken round-robin from the global order
    shards = [list(S.RankShard(order, 2, r, 2)) for r in range(2)]
    assert shards[0] == order[0:2] + order[4:6] and shards[1] == order[2:4] + order[6:8]
    assert len(S.RankShard(list(range(11)), 2, 0, 2)) == 4          # ragged tail dropped so both ranks run the same number of steps


def test_expand2square_and_process_images(env):
    from PIL import Image
    from llavamod.mm_utils import expand2square, get_model_name_from_path, process_images
    im = Image.new("RGB", (10, 4), (1, 2, 3))
    sq = expand2square(im, (9, 9, 9))
    assert sq.size == (10, 10) and sq.getpixel((0, 0)) == (9, 9, 9) and sq.getpixel((0, 3)) == (1, 2, 3) and sq.getpixel((0, 7)) == (9, 9, 9)
    tall = expand2square(Image.new("RGB", (4, 10), (1, 2, 3)), (9, 9, 9))
    assert tall.size == (10, 10) and tall.getpixel((2, 0)) == (9, 9, 9) and tall.getpixel((3, 0)) == (1, 2, 3)
    assert expand2square(sq, (0, 0, 0)) is sq
    proc = env.args("pad", "data_sft.json").image_processor
    out = process_images([im, sq], proc, types.SimpleNamespace(image_aspect_ratio="pad"))
    assert out.shape == (2, 3, 32, 32)
    assert process_images([im], proc, types.SimpleNamespace(image_aspect_ratio=None)).shape == (1, 3, 32, 32)
    assert get_model_name_from_path("/a/b/run1/checkpoint-200/") == "run1_checkpoint-200" and get_model_name_from_path("x/y") == "y"
