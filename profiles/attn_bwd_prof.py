import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "llava-mod_b200"))
from llavamod import kernels as K
B, T, nh, hd = 1, 2048, 16, int(os.environ.get("HD", "64"))
qkv = torch.randn(B * T, 3 * nh * hd, device="cuda").to(torch.bfloat16)
out, lse = K.attention_fwd(qkv, B, T, nh, nh, hd, True, need_lse=True)
go = torch.randn_like(out)
for _ in range(2):
    K.attention_bwd(qkv, out, go, lse, B, T, nh, nh, hd, True, hd ** -0.5)
torch.cuda.synchronize()
