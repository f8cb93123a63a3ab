"""The entry points RUN (not only parse): `align_train.train(argv)` and `dpo_train.train(argv)` with the flags of the reference's
shells (shells/train/qwen/dense2sparse_distillation.sh:48-88, preference_distillation.sh:48-88), from a saved tiny checkpoint directory
+ tokenizer + the golden JSON records, through the training loop, a mid-run checkpoint, auto-resume and the final save
(align_train.py:601-631: config.json + pytorch_model.bin = the full state dict with the reference's key names)."""
import json
import os
import types

import pytest
import torch

pytestmark = pytest.mark.gpu

from tests import helpers as Hh  # noqa: E402


def _save_tiny_checkpoints(root, golden_dir):
    """policy (dense tiny Qwen-1.5 shape, up-cycled to MoE by --policy_model_type sparse), teacher, CLIP tower dir, tokenizer files."""
    from llavamod.model import synthetic as S
    from tests.golden.make_data_golden import load_tokenizer
    tok = load_tokenizer(os.path.join(golden_dir, "tiny_tokenizer.json"))
    vocab = 424                                           # >= len(tok) = 420, multiple of 8
    teacher = S.make_teacher(dict(S.ARCH["tiny"], vocab_size=vocab, intermediate_size=320), "tiny", seed=0)
    policy = S.make_teacher(dict(S.ARCH["tiny"], vocab_size=vocab), "tiny", seed=1)
    policy.get_image_tower().load_state_dict(teacher.get_image_tower().state_dict())
    tower_dir = os.path.join(root, "openai-clip-tiny")
    os.makedirs(tower_dir)
    with open(os.path.join(tower_dir, "config.json"), "w") as f:
        json.dump(teacher.get_image_tower().config.to_dict(), f)
    torch.save({k: v.detach().cpu() for k, v in teacher.get_image_tower().image_tower.state_dict().items()}, os.path.join(tower_dir, "pytorch_model.bin"))
    dirs = {}
    for name, m in (("tiny-qwen1.5-teacher", teacher), ("tiny-qwen1.5-policy", policy)):
        d = os.path.join(root, name)
        m.config.mm_image_tower = tower_dir
        m.save_pretrained(d)
        tok.save_pretrained(d)
        dirs[name] = d
    return dirs["tiny-qwen1.5-policy"], dirs["tiny-qwen1.5-teacher"], tower_dir, tok


def _common_flags(policy, teacher, tower, out, golden_dir, data):
    return ["--policy_model_name_or_path", policy, "--ref_model_name_or_path", teacher, "--policy_model_type", "sparse", "--ref_model_type", "dense",
            "--moe_loss_enable", "True", "--moe_enable", "True", "--num_experts", "4", "--top_k_experts", "2", "--capacity_factor", "1.5",
            "--moe_mode", "sparse", "--use_residual", "False", "--router_aux_loss_coef", "0.01",
            "--train_modules", "mlp.gate_proj", "mlp.up_proj", "mlp.down_proj", "wg",
            "--deepspeed", "llavamod/config/dpconfig/zero2_offload.json", "--version", "qwen",
            "--data_path", os.path.join(golden_dir, data), "--image_folder", os.path.join(golden_dir, "data_imgs"),
            "--image_tower", tower, "--image_projector_type", "mlp2x_gelu", "--mm_vision_select_layer", "-2", "--mm_use_im_start_end", "False",
            "--mm_use_im_patch_token", "False", "--image_aspect_ratio", "pad", "--group_by_modality_length", "False", "--bf16", "True",
            "--output_dir", out, "--num_train_epochs", "1", "--per_device_train_batch_size", "2", "--per_device_eval_batch_size", "4",
            "--gradient_accumulation_steps", "1", "--evaluation_strategy", "no", "--save_strategy", "steps", "--save_steps", "2",
            "--save_total_limit", "3", "--weight_decay", "0.", "--warmup_ratio", "0.03", "--lr_scheduler_type", "cosine",
            "--logging_steps", "1", "--tf32", "True", "--model_max_length", "128", "--gradient_checkpointing", "True",
            "--dataloader_num_workers", "0", "--lazy_preprocess", "True", "--report_to", "none", "--cache_dir", "./cache_dir"]


def _check_final_save(out, trainer, n_ckpt_steps, use_cache=True):
    sd = torch.load(os.path.join(out, "pytorch_model.bin"), map_location="cpu", weights_only=True)
    live = trainer.model.state_dict()
    assert sorted(sd) == sorted(live)                                     # align_train.py:623-629: the FULL state dict
    for k in ("model.embed_tokens.weight", "model.layers.0.self_attn.q_proj.bias", "model.layers.0.mlp.deepspeed_moe.gate.wg.weight",
              "model.layers.0.mlp.deepspeed_moe.experts.deepspeed_experts.3.down_proj.weight", "model.layers.1.mlp.gate_proj.weight",
              "model.norm.weight", "lm_head.weight", "model.mm_projector.image_spatial_proj.0.weight",
              "model.image_tower.image_tower.vision_model.pre_layrnorm.weight"):
        assert k in sd, k
    assert sd["model.layers.0.mlp.deepspeed_moe.gate.wg.weight"].dtype == torch.float32
    for k, v in sd.items():
        assert torch.equal(v, live[k].detach().cpu()), k
    cfg = json.load(open(os.path.join(out, "config.json")))
    # align_train.py:621 flips use_cache back on before the final save; dpo_train.py:560 has that line commented out
    assert cfg["model_type"] == "moe_llava_qwen1_5" and cfg["moe"]["num_experts"] == [4] and cfg["use_cache"] is use_cache
    for s in n_ckpt_steps:
        d = os.path.join(out, "checkpoint-%d" % s)
        assert sorted(os.listdir(d)) == ["config.json", "optimizer.pt", "pytorch_model.bin", "rng_state.pth", "trainer_state.json"]


def test_align_train_entry_point_runs_to_the_final_save(tmp_path, golden_dir):
    from llavamod.train import align_train
    policy, teacher, tower, tok = _save_tiny_checkpoints(str(tmp_path), golden_dir)
    out = str(tmp_path / "out_mimic")
    argv = _common_flags(policy, teacher, tower, out, golden_dir, "data_sft.json") + ["--loss_type", "kd_lm", "--learning_rate", "1e-3",
                                                                                         "--num_train_epochs", "2"]
    policy_before = torch.load(os.path.join(policy, "pytorch_model.bin"), map_location="cpu", weights_only=True)
    tr = align_train.train(argv)
    assert tr.state.global_step == 4                                      # 4 records / batch 2 = 2 steps per epoch, 2 epochs
    logs = [json.loads(l) for l in open(os.path.join(out, "trainer_log.jsonl"))]
    assert [l["step"] for l in logs] == [1, 2, 3, 4]
    for l in logs:
        assert all(k in l for k in ("loss", "learning_rate", "epoch", "loss/align", "loss/lm", "loss/moe_balance")) and l["loss"] == l["loss"]
    _check_final_save(out, tr, (2, 4))
    sd = torch.load(os.path.join(out, "pytorch_model.bin"), map_location="cpu", weights_only=True)
    # --train_modules: FFN experts + router moved, attention / embeddings / tower did not
    assert not torch.equal(sd["model.layers.0.mlp.deepspeed_moe.experts.deepspeed_experts.0.gate_proj.weight"], policy_before["model.layers.0.mlp.gate_proj.weight"])
    assert torch.equal(sd["model.layers.0.self_attn.q_proj.weight"], policy_before["model.layers.0.self_attn.q_proj.weight"])
    assert torch.equal(sd["model.embed_tokens.weight"], policy_before["model.embed_tokens.weight"])
    # a second launch finds checkpoint-4 (align_train.py:601-604), resumes at the end of training and rewrites the same final files
    tr2 = align_train.train(argv)
    assert tr2.state.global_step == 4 and len(tr2.state.log_history) == 0
    sd2 = torch.load(os.path.join(out, "pytorch_model.bin"), map_location="cpu", weights_only=True)
    for k in sd:
        assert torch.equal(sd[k], sd2[k]), k


def test_dpo_train_entry_point_runs_to_the_final_save(tmp_path, golden_dir):
    from llavamod.train import dpo_train
    policy, teacher, tower, tok = _save_tiny_checkpoints(str(tmp_path), golden_dir)
    out = str(tmp_path / "out_pref")
    argv = _common_flags(policy, teacher, tower, out, golden_dir, "data_dpo.json") + ["--loss_type", "sigmoid", "--learning_rate", "2e-4",
                                                                                         "--num_train_epochs", "2", "--save_steps", "1"]
    tr = dpo_train.train(argv)
    assert tr.state.global_step == 2                                      # 2 preference records / batch 2 = 1 step per epoch
    logs = [json.loads(l) for l in open(os.path.join(out, "trainer_log.jsonl"))]
    for l in logs:
        for k in ("loss", "loss/reward", "loss/moe_balance", "loss/policy_chosen", "rewards/chosen", "rewards/rejected", "rewards/accuracies",
                  "rewards/margins", "logps/chosen", "logps/rejected"):
            assert k in l and l[k] == l[k], k
    _check_final_save(out, tr, (1, 2), use_cache=False)


def test_resumed_run_reproduces_the_uninterrupted_loss_sequence(tmp_path, golden_dir):
    """N2: stop after step 2 of 5 (checkpoint-2), start a NEW process-like trainer from the checkpoint: it skips the two consumed batches
    of the epoch, restores optimizer arenas, LR schedule position and the RNG streams (router noise), and logs exactly the losses the
    uninterrupted run logged for steps 3-5 (bit-for-bit: same kernels, same inputs, same noise)."""
    from llavamod.train import align_train
    policy, teacher, tower, tok = _save_tiny_checkpoints(str(tmp_path), golden_dir)

    def run(out, max_steps):
        argv = _common_flags(policy, teacher, tower, out, golden_dir, "data_sft.json") + [
            "--loss_type", "kd_lm", "--learning_rate", "1e-3", "--per_device_train_batch_size", "1", "--save_steps", "2", "--max_steps", str(max_steps),
            "--seed", "7"]
        os.environ["LLAVAMOD_CUDA_GRAPHS"] = "0"
        try:
            return align_train.train(argv)
        finally:
            os.environ.pop("LLAVAMOD_CUDA_GRAPHS", None)

    full = run(str(tmp_path / "a"), 5)
    want = {h["step"]: (h["loss"], h["loss/align"], h["learning_rate"]) for h in full.state.log_history}
    assert sorted(want) == [1, 2, 3, 4, 5]
    import shutil
    os.makedirs(tmp_path / "b")
    shutil.copytree(tmp_path / "a" / "checkpoint-2", tmp_path / "b" / "checkpoint-2")
    resumed = run(str(tmp_path / "b"), 5)
    got = {h["step"]: (h["loss"], h["loss/align"], h["learning_rate"]) for h in resumed.state.log_history}
    assert sorted(got) == [3, 4, 5]
    for s in (3, 4, 5):
        assert got[s][2] == want[s][2], (s, got[s], want[s])                         # learning rate: same schedule position
        for x, y in zip(got[s][:2], want[s][:2]):
            assert abs(x - y) <= 1e-5 * abs(y), (s, got[s], want[s])                 # same loss sequence (fp32 atomics order the only noise)
    a = torch.load(tmp_path / "a" / "pytorch_model.bin", map_location="cpu", weights_only=True)
    b = torch.load(tmp_path / "b" / "pytorch_model.bin", map_location="cpu", weights_only=True)
    for k in a:          # a few kernels reduce with fp32 atomics (global gradient norm, dQ, split-K): two runs agree to rounding, not bit for bit
        assert (a[k].float() - b[k].float()).norm() <= 1e-4 * a[k].float().norm() + 1e-9, k
    # an interrupted run that is NOT resumed properly (fresh optimizer state, data from the start) is visibly different: the check has teeth
    assert any(abs(want[s][0] - want[2][0]) > 1e-4 for s in (3, 4, 5))
