import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "llava-mod_b200")); sys.path.insert(0, ROOT)
from llavamod import kernels as K
from tests.test_attn_gpu import ref_attn
B, T, nh, nkv, hd, causal = [int(x) for x in sys.argv[1:7]]
g = torch.Generator(device="cuda").manual_seed(T + hd)
qkv = torch.randn(B * T, (nh + 2 * nkv) * hd, device="cuda", generator=g).to(torch.bfloat16).requires_grad_(True)
out = K.AttnFn.apply(qkv, B, T, nh, nkv, hd, bool(causal), None)
go = torch.randn(B * T, nh * hd, device="cuda", generator=g).to(torch.bfloat16)
torch.cuda.synchronize(); print("fwd ok", flush=True)
out.backward(go)
torch.cuda.synchronize(); print("bwd ran", flush=True)
x = qkv.detach().float().requires_grad_(True)
ref, _ = ref_attn(x, B, T, nh, nkv, hd, bool(causal), hd ** -0.5)
ref.backward(go.float())
for name, sl in (("dq", slice(0, nh * hd)), ("dk", slice(nh * hd, (nh + nkv) * hd)), ("dv", slice((nh + nkv) * hd, None))):
    a, r = qkv.grad[:, sl].float(), x.grad[:, sl]
    print(name, "rel err", (a - r).norm().item() / r.norm().item(), flush=True)
# error map of dq per (64-query block, head)
a, r = qkv.grad[:, : nh * hd].float().view(B, T, nh, hd), x.grad[:, : nh * hd].view(B, T, nh, hd)
for bb in range(B):
    for h in range(nh):
        errs = []
        for q0 in range(0, T, 64):
            d = (a[bb, q0:q0 + 64, h] - r[bb, q0:q0 + 64, h]).norm().item() / (r[bb, q0:q0 + 64, h].norm().item() + 1e-9)
            errs.append("%.3f" % d)
        print("b", bb, "h", h, " ".join(errs))
