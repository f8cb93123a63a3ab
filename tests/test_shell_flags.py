"""Every flag of the reference's six Qwen training shells (shells/train/qwen/*.sh) must be accepted by the matching entry point's
argument dataclasses (SURVEY section 8b "Entry points").  The shells are read from the reference tree, so this runs in the build
container only."""
import os
import re
import shlex

import pytest

REF = os.environ.get("LLAVAMOD_REFERENCE", "/root/reference")
SHELLS = os.path.join(REF, "shells", "train", "qwen")
pytestmark = pytest.mark.skipif(not os.path.isdir(SHELLS), reason="reference tree not present (GPU box)")

ENTRY = {"pretrain.sh": "train", "finetune.sh": "train", "finetune_moe.sh": "train", "dense2dense_distillation.sh": "align",
         "dense2sparse_distillation.sh": "align", "preference_distillation.sh": "dpo"}


def shell_argv(path):
    text = open(path).read()
    env = {}
    for m in re.finditer(r"^([A-Z_][A-Z0-9_]*)=(.*)$", text, re.M):
        if "deepspeed" in m.group(2):                             # the launch line itself starts with VAR=1 VAR=1 deepspeed ...
            continue
        val = shlex.split(m.group(2).split("#")[0])
        env[m.group(1)] = val[0] if val else ""
    cmd = re.sub(r"\\[ \t]*\n", " ", text[text.index("deepspeed llavamod/train/"):])
    cmd = re.sub(r"\$\{(\w+)\}", lambda m: env.get(m.group(1), "x"), cmd)
    toks = shlex.split(cmd)
    return toks[2:], toks[1]                                      # drop "deepspeed <script>"


@pytest.mark.parametrize("shell", sorted(ENTRY))
def test_shell_flags_parse(shell):
    from llavamod.config.args import (AlignArguments, DataArguments, DPOArguments, ModelArguments, TrainingArguments,
                                      parse_args_into_dataclasses)
    argv, script = shell_argv(os.path.join(SHELLS, shell))
    kind = ENTRY[shell]
    assert script.endswith({"train": "train.py", "align": "align_train.py", "dpo": "dpo_train.py"}[kind])
    classes = {"train": (ModelArguments, DataArguments, TrainingArguments),
               "align": (ModelArguments, DataArguments, TrainingArguments, AlignArguments),
               "dpo": (ModelArguments, DataArguments, TrainingArguments, DPOArguments)}[kind]
    out = parse_args_into_dataclasses(classes, argv)
    m, d, t = out[:3]
    assert t.output_dir and t.model_max_length >= 1024 and t.bf16 and d.data_path and d.image_folder
    if shell == "pretrain.sh":
        assert m.tune_mm_mlp_adapter and t.learning_rate == 1e-3
    if shell == "finetune_moe.sh":
        assert m.moe_enable and m.train_modules and m.num_experts
    if kind == "align":
        assert out[3].loss_type in ("kd_lm", "only_kd") and out[3].policy_model_type in ("dense", "sparse")
    if kind == "dpo":
        assert out[3].loss_type in ("sigmoid", "hinge", "ipo", "kto_pair")
    assert os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "llava-mod_b200", "llavamod", "train",
                                       os.path.basename(script)))
