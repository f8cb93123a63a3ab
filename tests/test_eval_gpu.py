"""SURVEY section 8f row N4 -- eval path: `generate` (the reference's eval loop runs HF generate with use_cache=False,
eval/model_vqa_loader.py:119-130) against a greedy loop over the CPU oracle, the Eval classes rebuilt from a saved checkpoint, and
`load_pretrained_model`."""
import os
import types

import pytest
import torch

pytestmark = pytest.mark.gpu

from tests import helpers as Hh  # noqa: E402


def _oracle_greedy(model, batch, steps):
    """Greedy decoding with the oracle: full forward each step (no cache), argmax of the last position.  Also returns the top-2 margin so the
    test can tell a genuine disagreement from a bf16 near-tie."""
    ids = batch["input_ids"].clone()
    toks, margins = [], []
    sd = Hh.oracle_state(model)
    for _ in range(steps):
        b = dict(batch, input_ids=ids, attention_mask=torch.ones_like(ids, dtype=torch.bool), labels=torch.full_like(ids, -100))
        out, _ = Hh.oracle_forward(model, b, None, sd=sd)
        last = out["logits"][:, -1, :].float()
        top = last.topk(2, dim=-1)
        toks.append(top.indices[:, 0])
        margins.append((top.values[:, 0] - top.values[:, 1]))
        ids = torch.cat([ids, top.indices[:, :1]], dim=1)
    return torch.stack(toks, 1), torch.stack(margins, 1), ids


def test_greedy_generate_matches_oracle_loop():
    from llavamod.model import synthetic as S
    model = S.make_teacher(dict(S.ARCH["tiny"]), "tiny", seed=21)
    batch, _ = Hh.tiny_batch(model, B=2, Tt=24, seed=22)
    steps = 6
    ref_tok, margin, ref_ids = _oracle_greedy(model, batch, steps)
    out = model.generate(batch["input_ids"], images=batch["images"], max_new_tokens=steps, do_sample=False, use_cache=False)
    assert out.shape == (2, 24 + steps) and torch.equal(out[:, :24].cpu(), batch["input_ids"])      # prompt ids (with -200) come back first
    got = out[:, 24:].cpu()
    for b in range(2):
        for s in range(steps):
            if got[b, s] != ref_tok[b, s]:
                assert margin[b, s] < 5e-2, (b, s, int(got[b, s]), int(ref_tok[b, s]), float(margin[b, s]))   # only a bf16 near-tie may differ
                break                                                                                            # sequences diverge afterwards
    assert torch.equal(got[:, 0], ref_tok[:, 0]) or float(margin[:, 0].min()) < 5e-2


def test_sampling_eos_and_stopping_criteria():
    from llavamod.model import synthetic as S
    model = S.make_teacher(dict(S.ARCH["tiny"]), "tiny", seed=23)
    batch, _ = Hh.tiny_batch(model, B=1, Tt=20, seed=24)
    g = torch.Generator(device="cuda").manual_seed(0)
    a = model.generate(batch["input_ids"], images=batch["images"], max_new_tokens=5, do_sample=True, temperature=0.7, top_p=0.9, generator=g)
    g = torch.Generator(device="cuda").manual_seed(0)
    b = model.generate(batch["input_ids"], images=batch["images"], max_new_tokens=5, do_sample=True, temperature=0.7, top_p=0.9, generator=g)
    assert torch.equal(a, b) and a.shape[1] == 25
    first = int(model.generate(batch["input_ids"], images=batch["images"], max_new_tokens=1)[0, -1])
    stopped = model.generate(batch["input_ids"], images=batch["images"], max_new_tokens=8, eos_token_id=first)
    assert stopped.shape[1] == 21                                       # the very first token is EOS
    calls = []
    crit = lambda ids, scores: (calls.append(ids.shape[1]) or ids.shape[1] >= 23)      # noqa: E731
    out = model.generate(batch["input_ids"], images=batch["images"], max_new_tokens=8, stopping_criteria=[crit])
    assert out.shape[1] == 23 and calls == [21, 22, 23]
    model.resize_token_embeddings(100)                                  # narrowed vocabulary: nothing >= 100 may be produced
    out = model.generate(batch["input_ids"], images=batch["images"], max_new_tokens=6)
    assert int(out[0, 20:].max()) < 100
    with pytest.raises(NotImplementedError):
        model.generate(batch["input_ids"], images=batch["images"], num_beams=3)


def test_eval_moe_class_from_saved_checkpoint_and_builder(tmp_path, golden_dir):
    """Train-side sparse student -> save_pretrained -> EvalLLaVAMoD...ForCausalLM.from_pretrained (experts rebuilt from config.moe) ->
    same logits; load_pretrained_model picks the class from the directory name and wires tokenizer / processor."""
    from llavamod.model import EvalLLaVAMoDQwen1_5ForCausalLM
    from llavamod.model.builder import load_pretrained_model, pick_eval_class
    from tests.golden.make_data_golden import load_tokenizer
    student, _ = Hh.tiny_pair(vocab=424)
    path = os.path.join(tmp_path, "llava-qwen1.5-tiny-moe")
    student.config.mm_image_tower = dict(student.get_image_tower().config.to_dict())
    student.save_pretrained(path)
    assert pick_eval_class("LLaVA-Qwen1.5-tiny-MoE") is EvalLLaVAMoDQwen1_5ForCausalLM
    tok = load_tokenizer(os.path.join(golden_dir, "tiny_tokenizer.json"))
    tokenizer, model, processor, ctx = load_pretrained_model(path, None, "llava-qwen1.5-tiny-moe", tokenizer=tok)
    assert isinstance(model, EvalLLaVAMoDQwen1_5ForCausalLM) and not model.training and ctx == 2048
    assert processor["image"] is not None and model._active_vocab == len(tokenizer)
    batch, noise = Hh.tiny_batch(student, B=1, Tt=20, seed=31)
    student.eval()
    with torch.no_grad():
        a = student(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"], images=batch["images"], moe_noise=[n.cuda() for n in noise])
        b = model(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"], images=batch["images"], moe_noise=[n.cuda() for n in noise])
    assert torch.equal(a.logits, b.logits)
    out = model.generate(batch["input_ids"], images=batch["images"], max_new_tokens=3)
    assert out.shape == (1, 23) and int(out[0, 20:].max()) < len(tokenizer)
