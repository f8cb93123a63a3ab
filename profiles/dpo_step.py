"""Preference-distillation step (BASELINE config 4 shapes on ONE GPU): 0.5B-4E policy <- 7B reference, chosen + rejected of T'=2048.
python profiles/dpo_step.py [steps]   -> ms per optimizer micro-step (CUDA events), loss, peak memory."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "llava-mod_b200"))
from llavamod.model import synthetic as S  # noqa: E402
from tests.helpers import make_trainer  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
t0 = time.time()
teacher = S.make_teacher("qwen1.5-7b", "clip-l-336", seed=0)
student = S.make_student("qwen1.5-0.5b", "clip-l-336", seed=1, margs=S.moe_args(train_modules=S.TRAIN_MODULES + ["deepspeed_experts"]),
                         share_tower_with=teacher)
tr = make_trainer(student, teacher, "sigmoid", kind="dpo", moe_loss_enable=True)
V, Tt = student.config.vocab_size, 2048 - 576 + 1
g = torch.Generator().manual_seed(3)


def batch():
    ch = torch.randint(0, V, (1, Tt), generator=g)
    ch[0, 5] = -200
    rj = ch.clone()
    k = int(0.4 * Tt)
    rj[0, k:] = torch.randint(0, V, (Tt - k,), generator=g)
    lab_c, lab_r = ch.clone(), rj.clone()
    lab_c[0, :k] = -100
    lab_r[0, :k] = -100
    m = torch.ones(1, Tt, dtype=torch.bool)
    return dict(chosen_input_ids=ch, chosen_labels=lab_c, chosen_attention_mask=m, rejected_input_ids=rj, rejected_labels=lab_r,
                rejected_attention_mask=m, images=[torch.randn(3, 336, 336, generator=g).to(torch.bfloat16)])


print("models built in %.1f s" % (time.time() - t0), flush=True)
for _ in range(2):
    loss = tr.training_step(student, batch())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
bs = [batch() for _ in range(steps)]
e0.record()
for b in bs:
    loss = tr.training_step(student, b)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
# 2 reference + 2 policy forwards (+ 2 policy backwards) of T'=2048: 2*30.17 + 2*7.6 + CLIP 0.38 TFLOP (SURVEY 8d figures)
print("dpo step: %.1f ms per pair  (%.2f pairs/s, ~%.0f nominal TFLOP/s), loss %.4f, peak memory %.1f GB"
      % (ms, 1e3 / ms, (2 * 30.17 + 2 * 7.6 + 0.38) / (ms * 1e-3), float(loss), torch.cuda.max_memory_allocated() / 2 ** 30), flush=True)
