"""Host-side logic of the widened rows (no GPU): which parameters each stage trains (train.py:476-486 / initialize_moe_modules),
eval-class dispatch by checkpoint name (builder.py:370-392), conversation rendering, and the per-rank sharding of the length-grouped order."""
import types

import pytest
import torch


def _tiny(kind):
    from llavamod.model import synthetic as S
    if kind == "dense":
        return S.make_teacher(dict(S.ARCH["tiny"]), "tiny", device="cpu", dtype=torch.float32)
    return S.make_student(dict(S.ARCH["tiny"]), "tiny", device="cpu", dtype=torch.float32, margs=S.moe_args())


def test_stage_parameter_selection():
    from llavamod.train.train import select_trainable
    targs = types.SimpleNamespace(tune_mm_mlp_adapter=False, freeze_mm_mlp_adapter=False)
    m = _tiny("dense")
    select_trainable(m, types.SimpleNamespace(tune_mm_mlp_adapter=True, moe_enable=False), targs)
    names = {n for n, p in m.named_parameters() if p.requires_grad}
    assert names and all("mm_projector" in n for n in names) and m.config.tune_mm_mlp_adapter and targs.tune_mm_mlp_adapter
    m = _tiny("dense")
    select_trainable(m, types.SimpleNamespace(tune_mm_mlp_adapter=False, moe_enable=False), targs)
    names = {n for n, p in m.named_parameters() if p.requires_grad}
    assert any("embed_tokens" in n for n in names) and any("norm" in n for n in names) and not any("image_tower" in n for n in names)
    targs.freeze_mm_mlp_adapter = True
    select_trainable(m, types.SimpleNamespace(tune_mm_mlp_adapter=False, moe_enable=False), targs)
    assert not any("mm_projector" in n for n, p in m.named_parameters() if p.requires_grad)
    targs.freeze_mm_mlp_adapter = False
    s = _tiny("sparse")                                   # initialize_moe_modules already froze everything but --train_modules
    before = {n for n, p in s.named_parameters() if p.requires_grad}
    select_trainable(s, types.SimpleNamespace(tune_mm_mlp_adapter=False, moe_enable=True), targs)
    assert {n for n, p in s.named_parameters() if p.requires_grad} == before
    assert all(("mlp" in n or "wg" in n or "mm_projector" in n) for n in before)


def test_eval_class_dispatch_by_name():
    from llavamod import model as M
    from llavamod.model.builder import pick_eval_class
    assert pick_eval_class("LLaVA-MoD-Qwen1.5-0.5B-moe") is M.EvalLLaVAMoDQwen1_5ForCausalLM
    assert pick_eval_class("llava-qwen-1.5-1.8b") is M.LlavaQwen1_5ForCausalLM
    assert pick_eval_class("llavaqwen2-0.5b-MoE") is M.EvalLLaVAMoDQwen2ForCausalLM
    assert pick_eval_class("llava-qwen-2-7b") is M.LlavaQwen2ForCausalLM
    for bad in ("llava-qwen-7b", "llava-phi2-moe"):
        with pytest.raises(NotImplementedError):
            pick_eval_class(bad)


def test_conversation_rendering():
    from llavamod import conversation as C
    c = C.conv_templates["qwen"].copy()
    c.append_message(c.roles[0], "<image>\nWhat is this?")
    c.append_message(c.roles[1], "A bird.")
    c.append_message(c.roles[0], "Sure?")
    c.append_message(c.roles[1], None)
    assert c.get_prompt() == (C._CHAT_SYSTEM + " USER: <image>\nWhat is this? ASSISTANT: A bird.<|endoftext|>USER: Sure? ASSISTANT:")
    assert C.conv_templates["qwen"].messages == []         # copy() does not alias the template's message list
    p = C.conv_templates["plain"].copy()
    p.append_message("", "<image>")
    p.append_message("", "a caption")
    assert p.get_prompt() == "<image>\na caption\n"
    assert C.set_default_conversation("plain") is C.conv_llava_plain and C.set_default_conversation("qwen") is C.conv_phi


def test_trainer_uses_the_length_grouped_sampler_per_rank():
    from llavamod.train.trainer_base import BaseTrainer

    class DS(torch.utils.data.Dataset):
        modality_lengths = [5, 9, -3, 12, 7, -8, 2, 11, 6, 4, 10, 1, -13, 14, 15, 16]

        def __len__(self):
            return 16

        def __getitem__(self, i):
            return i

    args = types.SimpleNamespace(group_by_modality_length=True, per_device_train_batch_size=2, gradient_accumulation_steps=2,
                                 dataloader_num_workers=0, seed=0)
    seen = []
    for rank in range(2):
        tr = BaseTrainer.__new__(BaseTrainer)
        tr.args, tr.train_dataset, tr.data_collator, tr.world_size, tr.rank = args, DS(), (lambda b: list(b)), 2, rank
        torch.manual_seed(0)
        seen.append([i for batch in tr.get_train_dataloader() for i in batch])
    assert len(seen[0]) == len(seen[1]) == 8 and not set(seen[0]) & set(seen[1])
    assert sorted(seen[0] + seen[1]) == list(range(16))


def test_load_tokenizer_from_a_local_directory(golden_dir, tmp_path):
    """align_train.py:360-369,436-441: slow-or-fast AutoTokenizer from the checkpoint directory, right padding, `<|extra_0|>` as unk, pad = unk,
    conversation template from --version."""
    import os
    from llavamod import conversation as C
    from llavamod.train.align_train import load_tokenizer
    from tests.golden.make_data_golden import load_tokenizer as golden_tokenizer
    golden_tokenizer(os.path.join(golden_dir, "tiny_tokenizer.json")).save_pretrained(str(tmp_path))
    tok = load_tokenizer(types.SimpleNamespace(version="qwen"), types.SimpleNamespace(cache_dir=None, model_max_length=64), str(tmp_path))
    assert tok.unk_token == "<|extra_0|>" and tok.pad_token == tok.unk_token and tok.padding_side == "right" and tok.model_max_length == 64
    assert C.default_conversation is C.conv_phi
    with pytest.raises(NotImplementedError):
        load_tokenizer(types.SimpleNamespace(version="llama_2"), types.SimpleNamespace(cache_dir=None, model_max_length=64), str(tmp_path))
    C.set_default_conversation("qwen")
