"""GEMM timing experiments (not a test): 1-CTA vs 2-CTA kernel, with and without epilogue stores, a few shapes."""
import os, sys, subprocess
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "llava-mod_b200"))
from llavamod import kernels as K
shapes = [(2048, 22016, 4096), (2048, 4096, 4096), (2048, 4096, 11008), (2048, 12288, 4096), (8192, 8192, 8192)]
for (M, N, Kd) in shapes:
    a = torch.randn(M, Kd, device="cuda").to(torch.bfloat16); b = torch.randn(N, Kd, device="cuda").to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    res = {}
    for name, fn in (("lmod", lambda: K.gemm(a, b, out=out)), ("cublas", lambda: torch.mm(a, b.t(), out=out))):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        res[name] = 2.0 * M * N * Kd * 20 / (e0.elapsed_time(e1) * 1e-3) / 1e12
    print("%s 2CTA=%s NOSTORE=%s : %5dx%5dx%5d lmod %.0f cublas %.0f TFLOP/s" % (os.environ.get("TAG", ""), os.environ.get("LMOD_GEMM_2CTA", "1"), os.environ.get("LMOD_GEMM_NOSTORE", "0"), M, N, Kd, res["lmod"], res["cublas"]))
