"""Shared fixtures for the GPU parity tests, smoke() and bench.py: tiny config-1 models, seeded batches, and the
oracle-side evaluation of the same step."""
import types

import torch

from oracle import restated as R


def tiny_pair(device="cuda", student_layers=2, teacher_layers=2, vocab=512, seed=0):
    from llavamod.model import synthetic as S
    arch_s = dict(S.ARCH["tiny"], num_hidden_layers=student_layers, vocab_size=vocab)
    arch_t = dict(S.ARCH["tiny"], num_hidden_layers=teacher_layers, vocab_size=vocab, intermediate_size=320)
    teacher = S.make_teacher(arch_t, "tiny", device=device, seed=seed)
    student = S.make_student(arch_s, "tiny", device=device, seed=seed + 1, margs=S.moe_args(), share_tower_with=teacher)
    return student, teacher


def tiny_batch(student, B=2, Tt=40, seed=0, pad=(0, 0), n_img_tokens=1):
    """config-1 shape: 32x32 image -> 16 patches, text length chosen so the spliced length is 40-1+16 = 55 (odd on purpose)."""
    g = torch.Generator().manual_seed(seed)
    V = student.config.vocab_size
    ids = torch.randint(0, V, (B, Tt), generator=g)
    ids[:, 5] = -200
    mask = torch.ones(B, Tt, dtype=torch.bool)
    for b, p in enumerate(pad[:B]):
        if p:
            mask[b, Tt - p:] = False
    labels = ids.clone()
    labels[:, : int(0.4 * Tt)] = -100
    labels[~mask] = -100
    images = [torch.randn(3, 32, 32, generator=g).to(torch.bfloat16) for _ in range(B)]
    Tn = Tt - 1 + 16
    n_moe = sum(1 for l in student.model.layers if hasattr(l.mlp, "deepspeed_moe"))
    E = 4
    noise = [R.gumbel_noise((B * Tn, E), g) for _ in range(n_moe)]
    return dict(input_ids=ids, labels=labels, attention_mask=mask, images=images), noise


def cfgs_of(model):
    c = model.config
    t = model.get_image_tower().config
    cc = R.ClipCfg(hidden=t.hidden_size, inter=t.intermediate_size, layers=t.num_hidden_layers, heads=t.num_attention_heads,
                   image=t.image_size, patch=t.patch_size, eps=t.layer_norm_eps, select_layer=c.mm_vision_select_layer)
    moe_layers = [i for i, l in enumerate(model.model.layers) if hasattr(l.mlp, "deepspeed_moe")]
    kw = {}
    if moe_layers:
        m = model.model.layers[moe_layers[0]].mlp
        kw = dict(moe_layers=moe_layers, num_experts=m.num_experts, capacity_factor=m.capacity_factor, min_capacity=m.min_capacity,
                  aux_coef=model.router_aux_loss_coef)
    lc = R.LMCfg(hidden=c.hidden_size, inter=c.intermediate_size, layers=c.num_hidden_layers, heads=c.num_attention_heads,
                 kv_heads=c.num_key_value_heads, vocab=c.vocab_size, rope_theta=c.rope_theta, eps=c.rms_norm_eps,
                 tie=bool(getattr(c, "tie_word_embeddings", False)), kd_vocab=min(R.KD_VOCAB, c.vocab_size), **kw)
    return lc, cc


def oracle_state(model, dtype=torch.float32):
    return {k: v.detach().to("cpu").to(dtype if v.dtype != torch.float32 or dtype == torch.float32 else v.dtype) for k, v in model.state_dict().items()}


def oracle_forward(model, batch, noise=None, sd=None, dtype=torch.float32):
    lc, cc = cfgs_of(model)
    sd = sd if sd is not None else oracle_state(model, dtype)
    imgs = [im.to(dtype) for im in batch["images"]]
    return R.llava_forward(sd, lc, cc, batch["input_ids"], batch["attention_mask"], batch["labels"], imgs, noise), lc


def oracle_mimic_loss(student, teacher, batch, noise, loss_type="kd_lm", moe_loss_enable=True, sd_s=None, sd_t=None):
    with torch.no_grad():
        t_out, _ = oracle_forward(teacher, batch, sd=sd_t)
    s_out, lc = oracle_forward(student, batch, noise, sd=sd_s)
    return R.mimic_compute_loss(s_out, t_out["logits"], loss_type, moe_loss_enable, False, lc.kd_vocab)


def make_trainer(student, teacher, loss_type="kd_lm", accum=1, lr=2e-5, max_steps=100, kind="align", moe_loss_enable=True):
    from llavamod.config.args import TrainingArguments
    from llavamod.train.align_trainer import AlignTrainer
    from llavamod.train.dpo_trainer import DPOTrainer
    args = TrainingArguments(output_dir="/tmp/lmod_out", per_device_train_batch_size=1, gradient_accumulation_steps=accum,
                             learning_rate=lr, weight_decay=0.0, warmup_ratio=0.03, lr_scheduler_type="cosine", max_steps=max_steps,
                             logging_steps=0, save_strategy="no", bf16=True)
    args.moe_enable = True
    cls = AlignTrainer if kind == "align" else DPOTrainer
    tr = cls(model=student, ref_model=teacher, args=args, loss_type=loss_type, moe_loss_enable=moe_loss_enable)
    tr._total_steps = max_steps
    return tr
