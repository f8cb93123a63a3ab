"""Do the GPU (bf16 activations) and the fp32 CPU oracle route the same tokens to the same experts?  Tiny config, the batches of the
preference-gradient test.  python profiles/route_diag.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "llava-mod_b200"))
from oracle import restated as R
from tests import helpers as Hh

student, teacher = Hh.tiny_pair()
for seed in (4, 7, 8, 21, 22, 23, 24, 25, 26):
    batch, noise = Hh.tiny_batch(student, seed=seed)
    lc, cc = Hh.cfgs_of(student)
    rec = []
    sd = Hh.oracle_state(student)
    R.llava_forward(sd, lc, cc, batch["input_ids"], batch["attention_mask"], batch["labels"], [im.float() for im in batch["images"]], noise, record=rec)
    with torch.no_grad():
        r = student.forward_hidden(input_ids=batch["input_ids"], labels=batch["labels"], attention_mask=batch["attention_mask"], images=batch["images"],
                                   moe_noise=[n.cuda() for n in noise])
    g = r["records"][0]
    o = rec[0]
    idx = g["idx"].cpu().long()
    keep = g["row"].cpu() >= 0
    d1 = int((idx[:, 0] != o["idx1"]).sum()); d2 = int((idx[:, 1] != o["idx2"]).sum())
    k1 = int((keep[:, 0] != o["keep1"]).sum()); k2 = int((keep[:, 1] != o["keep2"]).sum())
    gap = (o["logits"].topk(2, dim=1).values[:, 0] - o["logits"].topk(2, dim=1).values[:, 1]).min().item()
    print("seed %d: tokens %d  top1 differs %d  top2 differs %d  keep1 differs %d keep2 differs %d  drops %d  min top-2 logit gap %.4f  max |logit diff| %.4f"
          % (seed, idx.shape[0], d1, d2, k1, k2, int((~o["keep1"]).sum() + (~o["keep2"]).sum()), gap, (g["logits"].cpu() - o["logits"]).abs().max().item()))
