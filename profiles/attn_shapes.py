"""Attention forward over a few shapes, CUDA-event timed (50 launches back to back after warm-up): used to A/B LMOD_ATTN_SPLIT / LMOD_ATTN_OPT,
which the library reads once per process.   LMOD_ATTN_SPLIT=1 LMOD_ATTN_VERBOSE=1 python profiles/attn_shapes.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "llava-mod_b200"))
from llavamod._C import call, ptr

def run(B, T, nh, hd, causal):
    qkv = (torch.randn(B * T, 3 * nh * hd, device="cuda") * 0.5).to(torch.bfloat16)
    out = torch.empty(B * T, nh * hd, device="cuda", dtype=torch.bfloat16)
    f = lambda: call("lmod_attn_fwd", ptr(qkv), qkv.stride(0), B, T, nh, nh, hd, int(causal), hd ** -0.5, ptr(out), out.stride(0), None, None, None)
    for _ in range(5): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / 50
    fl = 4.0 * B * nh * T * T * hd * (0.5 if causal else 1.0)
    print("split=%s opt=%s B%d T%-5d nh%-3d hd%-4d causal=%d  %8.1f us  %7.1f TFLOP/s" % (os.environ.get("LMOD_ATTN_SPLIT", "auto"), os.environ.get("LMOD_ATTN_OPT", "dflt"), B, T, nh, hd, causal, us, fl / us / 1e6), flush=True)

for shape in ((1, 512, 37, 64, False), (1, 1024, 37, 64, False), (1, 2048, 37, 64, False), (1, 512, 37, 128, False), (1, 1024, 37, 128, False), (1, 2048, 16, 64, True), (1, 1024, 16, 64, True), (1, 4096, 16, 64, True), (1, 2048, 8, 128, True), (1, 2048, 32, 128, True), (1, 577, 16, 64, False), (2, 577, 16, 64, False), (4, 2048, 16, 64, True)):
    run(*shape)
