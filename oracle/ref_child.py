"""TEST INFRASTRUCTURE ONLY -- child process that runs the REFERENCE's dense path (see oracle/ref_shim.py)."""
import os
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim  # noqa: E402


def main(req_path, resp_path):
    rq = torch.load(req_path, weights_only=False)
    tmp = tempfile.mkdtemp()
    m = ref_shim.build_tiny_dense(tmp, **rq["kw"])
    if rq.get("padding_side"):
        m.config.tokenizer_padding_side = rq["padding_side"]
    out = {}
    if rq.get("clip_images") is not None:
        out["clip_features"] = m.get_model().get_image_tower()(rq["clip_images"]).detach()
    if rq.get("input_ids") is not None:
        r = m(input_ids=rq["input_ids"], labels=rq["labels"], attention_mask=rq["attention_mask"], images=rq["images"], return_dict=True)
        out.update(logits=r.logits.detach(), labels=r.labels, loss=r.loss.detach())
    out["state_dict"] = {k: v.detach().clone() for k, v in m.state_dict().items()}
    torch.save(out, resp_path)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
