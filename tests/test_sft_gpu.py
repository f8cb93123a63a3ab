"""SURVEY section 8f row N3 -- the plain-CE stages (LLaVATrainer): dense SFT, MoE fine-tuning and adaptor pre-training, GPU path against
the CPU oracle's own `.loss` (oracle/restated.py::llava_forward) stepped with the restated AdamW / cosine schedule / clipping."""
import os
import types

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import restated as R  # noqa: E402
from tests import helpers as Hh  # noqa: E402


def _sft_trainer(model, lr, steps, accum=1, **flags):
    from llavamod.config.args import TrainingArguments
    from llavamod.train.llava_trainer import LLaVATrainer
    args = TrainingArguments(output_dir=flags.pop("output_dir", "/tmp/lmod_sft"), per_device_train_batch_size=1, gradient_accumulation_steps=accum,
                             learning_rate=lr, weight_decay=0.0, warmup_ratio=0.03, lr_scheduler_type="cosine", max_steps=steps,
                             logging_steps=0, save_strategy="no", bf16=True)
    for k, v in flags.items():
        setattr(args, k, v)
    tr = LLaVATrainer(model=model, args=args)
    tr._total_steps = steps
    return tr


def _dense_student(seed=3):
    from llavamod.model import synthetic as S
    m = S.make_teacher(dict(S.ARCH["tiny"]), "tiny", seed=seed).train()
    for n, p in m.named_parameters():
        p.requires_grad = "image_tower" not in n
    return m


def _oracle_curve(model, keys, batches, noises, steps, lr):
    sd = Hh.oracle_state(model)
    params = [sd[k].requires_grad_(True) for k in keys]
    m = [torch.zeros_like(p) for p in params]
    v = [torch.zeros_like(p) for p in params]
    out = []
    for s in range(steps):
        o, _ = Hh.oracle_forward(model, batches[s], noises[s] or None, sd=sd)
        grads = [g.clone() for g in torch.autograd.grad(o["loss"], params)]
        R.clip_grad_norm(grads, 1.0)
        with torch.no_grad():
            R.adamw_step(params, grads, m, v, s + 1, R.cosine_lr(s, steps, lr))
        out.append(float(o["loss"]))
    return out


@pytest.mark.parametrize("kind", ["dense_full", "moe_finetune"])
def test_sft_loss_curve_tracks_oracle(kind):
    """Dense SFT trains every LM parameter (finetune.sh); MoE fine-tuning trains FFN experts + router (finetune_moe.sh) and its loss
    carries the auxiliary balance term (llava_qwen1_5_moe.py:431-434)."""
    if kind == "dense_full":
        model = _dense_student()
    else:
        model, _ = Hh.tiny_pair()
    steps, lr = 12, 1e-3
    pairs = [Hh.tiny_batch(model, seed=300 + s) for s in range(steps)]
    keys = [n for n, p in model.named_parameters() if p.requires_grad]
    ref = _oracle_curve(model, keys, [p[0] for p in pairs], [p[1] for p in pairs], steps, lr)
    tr = _sft_trainer(model, lr, steps)
    for s, (batch, noise) in enumerate(pairs):
        extra = dict(moe_noise=[n.cuda() for n in noise]) if noise else {}
        loss = float(tr.training_step(model, dict(batch, **extra)))
        assert abs(loss - ref[s]) < 2e-2 * abs(ref[s]), (kind, s, loss, ref[s])


def test_adapter_pretraining_trains_only_the_projector_and_saves_it(tmp_path):
    """pretrain.sh: --tune_mm_mlp_adapter True: only mm_projector moves; the final save is config.json + mm_projector.bin, which
    initialize_vision_modules of the next stage loads through --pretrain_mm_mlp_adapter."""
    from llavamod.model import synthetic as S
    from llavamod.train.train import select_trainable
    from llavamod.train.train_utils import safe_save_model_for_hf_trainer
    model = S.make_teacher(dict(S.ARCH["tiny"]), "tiny", seed=5).train()
    margs = types.SimpleNamespace(tune_mm_mlp_adapter=True, moe_enable=False)
    targs = types.SimpleNamespace(tune_mm_mlp_adapter=False, freeze_mm_mlp_adapter=False)
    select_trainable(model, margs, targs)
    names = [n for n, p in model.named_parameters() if p.requires_grad]
    assert names and all("mm_projector" in n for n in names)
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    tr = _sft_trainer(model, 1e-2, 4, output_dir=str(tmp_path), tune_mm_mlp_adapter=True)
    losses = [float(tr.training_step(model, Hh.tiny_batch(model, seed=20)[0])) for _ in range(4)]
    assert losses[-1] < losses[0]                         # same batch four times: the projector alone fits it a little
    for n, p in model.named_parameters():
        moved = not torch.equal(p.detach(), before[n])
        assert moved == ("mm_projector" in n), n
    safe_save_model_for_hf_trainer(tr, str(tmp_path))
    saved = torch.load(os.path.join(tmp_path, "mm_projector.bin"))
    assert os.path.exists(os.path.join(tmp_path, "config.json")) and sorted(saved) == sorted(k for k in model.state_dict() if "mm_projector" in k)
    tr.state.global_step = 4
    tr._save_checkpoint(model, None)                      # adaptor-only checkpoint folder (llava_trainer.py:249-272)
    assert sorted(os.listdir(os.path.join(tmp_path, "checkpoint-4"))) == ["config.json", "mm_projector.bin"]


def test_sft_graph_replay_matches_eager():
    """Un-padded micro-batches are captured into a CUDA graph after two eager warm-ups; replayed losses equal the eager ones (lr = 0)."""
    model = _dense_student(seed=9)
    batches = [Hh.tiny_batch(model, seed=40 + i)[0] for i in range(3)]
    tr = _sft_trainer(model, 0.0, 100)
    tr.use_cuda_graphs = False
    eager = [float(tr.training_step(model, dict(batches[i % 3]))) for i in range(6)]
    tr2 = _sft_trainer(model, 0.0, 100)
    tr2.optimizer = tr.optimizer
    got = [float(tr2.training_step(model, dict(batches[i % 3]))) for i in range(6)]
    assert tr2._graphs and any("graph" in e for e in tr2._graphs.values())
    for a, b in zip(eager, got):
        assert abs(a - b) < 1e-5 * abs(a) + 1e-6, (eager, got)
