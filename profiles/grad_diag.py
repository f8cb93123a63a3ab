"""Per-parameter relative gradient error (GPU bf16 path vs fp32 CPU oracle autograd) for the mimic and the preference step at the tiny config.
python profiles/grad_diag.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "llava-mod_b200"))
from oracle import restated as R
from tests import helpers as Hh
from tests.test_model_gpu import _dpo_inputs


def report(tag, student, sd_s):
    for n, p in student.named_parameters():
        if p.requires_grad:
            g, r = p.grad.float().cpu(), sd_s[n].grad
            cos = torch.nn.functional.cosine_similarity(g.flatten(), r.flatten(), dim=0).item()
            print("%-8s %-70s rel %.4f  cos %.5f  |ref| %.3e" % (tag, n, (g - r).norm().item() / (r.norm().item() + 1e-12), cos, r.norm().item()))


for kind in ("mimic", "dpo_sigmoid", "dpo_nomoe"):
    student, teacher = Hh.tiny_pair()
    sd_s = Hh.oracle_state(student)
    keys = [n for n, p in student.named_parameters() if p.requires_grad]
    for k in keys:
        sd_s[k].requires_grad_(True)
    if kind == "mimic":
        batch, noise = Hh.tiny_batch(student, seed=4)
        ref_loss, _ = Hh.oracle_mimic_loss(student, teacher, batch, noise, "kd_lm", sd_s=sd_s)
        ref_loss.backward()
        tr = Hh.make_trainer(student, teacher, "kd_lm")
        tr.create_optimizer().zero_grad()
        tr.compute_loss(student, dict(batch, moe_noise=[n.cuda() for n in noise])).backward()
    else:
        moe = kind == "dpo_sigmoid"
        bc, nc, br, nr, inputs = _dpo_inputs(student)
        with torch.no_grad():
            tc, _ = Hh.oracle_forward(teacher, bc)
            trj, _ = Hh.oracle_forward(teacher, br)
        pc, _ = Hh.oracle_forward(student, bc, nc, sd=sd_s)
        pr, _ = Hh.oracle_forward(student, br, nr, sd=sd_s)
        ref_loss, _ = R.dpo_compute_loss(pc, pr, tc["logits"], tc["labels"], trj["logits"], trj["labels"], 0.1, "sigmoid", moe)
        ref_loss.backward()
        tr = Hh.make_trainer(student, teacher, "sigmoid", kind="dpo", moe_loss_enable=moe)
        tr.create_optimizer().zero_grad()
        tr.compute_loss(student, inputs).backward()
    torch.cuda.synchronize()
    report(kind, student, sd_s)
