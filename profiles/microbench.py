"""Stand-alone launches of the hot kernels at config-2 shapes, small enough in memory for `ncu --set full` to replay cheaply.
    ncu --set full --clock-control none --import-source on -k regex:'gemm_tcgen05|attn_fwd|kl_fused|moe_route' -o gpurun_out/kernels_rN python profiles/microbench.py
Also prints CUDA-event timings (TFLOP/s, GB/s) when run without ncu."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "llava-mod_b200"))
from llavamod import kernels as K  # noqa: E402

dev = "cuda"
REPS = int(os.environ.get("REPS", "1"))


def timed(name, fn, work, unit):
    """REPS back-to-back launches replayed from a CUDA graph (as the training step runs them): device time per call, no host launch /
    allocation cost in the number."""
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(REPS):
                fn()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / REPS
    print("%-46s %8.3f ms  %9.1f %s" % (name, ms, work / (ms * 1e-3) / (1e12 if unit == "TFLOP/s" else 1e9), unit))


_FLUSH = None


def timed_cold(name, fn, work, unit):
    """HBM-bound kernels: L2 (126 MB) is flushed before every repetition, each repetition timed on its own; median."""
    global _FLUSH
    if _FLUSH is None:
        _FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    fn()
    ts = []
    for _ in range(REPS):
        _FLUSH.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    print("%-46s %8.3f ms  %9.1f %s (L2 flushed)" % (name, ms, work / (ms * 1e-3) / 1e9, unit))


def main():
    T, V = 2048, 151936
    # teacher gate|up GEMM, student gate|up, wgrad, lm_head
    for (M, N, Kd, a_mn, b_mn, tag) in [(T, 22016, 4096, False, False, "teacher gate|up fwd"), (T, 5632, 1024, False, False, "student gate|up fwd"),
                                        (T, 1024, 5632, False, True, "student dgrad (B MN-major)"), (5632, 1024, T, True, True, "student wgrad (A,B MN-major)"),
                                        (T, V, 1024, False, False, "student lm_head")]:
        a = torch.randn((Kd, M) if a_mn else (M, Kd), device=dev).to(torch.bfloat16)
        b = torch.randn((Kd, N) if b_mn else (N, Kd), device=dev).to(torch.bfloat16)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        timed("gemm %s %dx%dx%d" % (tag, M, N, Kd), lambda: K.gemm(a, b, a_mn=a_mn, b_mn=b_mn, out=out), 2.0 * M * N * Kd, "TFLOP/s")
    # grouped expert GEMM on compact rows (4 experts, ~1024 rows each)
    offs = torch.tensor([0, 1024, 2176, 3072, 4096], dtype=torch.int32, device=dev)
    xp = torch.randn(4608, 1024, device=dev).to(torch.bfloat16)
    w = torch.randn(4, 5632, 1024, device=dev).to(torch.bfloat16)
    h1 = torch.zeros(4608, 5632, device=dev, dtype=torch.bfloat16)
    timed("grouped expert gemm 4x[~1024,5632,1024]", lambda: K.grouped_gemm(xp, w, h1, offs, 0), 2.0 * 4096 * 5632 * 1024, "TFLOP/s")
    # fused SwiGLU MLP input (gate|up GEMM + SwiGLU epilogue) and the silu-backward dgrad epilogue, against their unfused forms
    for (M, H, I, tag, save) in [(T, 4096, 11008, "teacher", False), (T, 1024, 2816, "student dense", True)]:
        x = torch.randn(M, H, device=dev).to(torch.bfloat16)
        wgu = (torch.randn(2 * I, H, device=dev) * 0.02).to(torch.bfloat16)
        wdn = (torch.randn(H, I, device=dev) * 0.02).to(torch.bfloat16)
        timed("gemm_swiglu %s %dx%dx%d (h1 saved: %s)" % (tag, M, 2 * I, H, save), lambda: K.gemm_swiglu(x, wgu, save), 2.0 * M * 2 * I * H, "TFLOP/s")
        timed("  unfused: gemm + silu_mul", lambda: K.silu_mul(K.gemm(x, wgu)), 2.0 * M * 2 * I * H, "TFLOP/s")
        if save:
            _, h1s = K.gemm_swiglu(x, wgu, True)
            dy = torch.randn(M, H, device=dev).to(torch.bfloat16)
            timed("gemm_silu_bwd %s %dx%dx%d" % (tag, M, I, H), lambda: K.gemm_silu_bwd(dy, wdn, h1s), 2.0 * M * I * H, "TFLOP/s")
            timed("  unfused: dgrad gemm + silu_mul_bwd", lambda: K.silu_mul_bwd(K.gemm(dy, wdn, b_mn=True), h1s), 2.0 * M * I * H, "TFLOP/s")
    from llavamod.model.language_model.qwen2_core import rope_tables
    for (M, H, nh, hd, tag) in [(T, 4096, 32, 128, "teacher"), (T, 1024, 16, 64, "student")]:
        x = torch.randn(M, H, device=dev).to(torch.bfloat16)
        w = (torch.randn(3 * nh * hd, H, device=dev) * 0.02).to(torch.bfloat16)
        b = torch.randn(3 * nh * hd, device=dev).to(torch.bfloat16)
        cos, sin = rope_tables(hd, 2048, 1e6, torch.bfloat16, dev)
        pos = torch.arange(M, device=dev)
        K.FUSE_ROPE = "1"
        timed("qkv GEMM + RoPE epilogue %s %dx%dx%d" % (tag, M, 3 * nh * hd, H), lambda: K.qkv_rope(x, w, b, cos, sin, pos, nh, nh, hd), 2.0 * M * 3 * nh * hd * H, "TFLOP/s")
        K.FUSE_ROPE = "0"
        timed("  unfused: gemm + rope", lambda: K.qkv_rope(x, w, b, cos, sin, pos, nh, nh, hd), 2.0 * M * 3 * nh * hd * H, "TFLOP/s")
        K.FUSE_ROPE = "auto"
    # residual add in the projection's epilogue + plain RMSNorm  vs  projection + RMSNorm that adds the residual (teacher o_proj / down_proj)
    for (M, N, Kd, tag) in [(T, 4096, 4096, "teacher o_proj"), (T, 4096, 11008, "teacher down_proj"), (577, 1024, 4096, "CLIP fc2")]:
        x = torch.randn(M, Kd, device=dev).to(torch.bfloat16)
        w = (torch.randn(N, Kd, device=dev) * 0.02).to(torch.bfloat16)
        res = torch.randn(M, N, device=dev).to(torch.bfloat16)
        nw = torch.ones(N, device=dev, dtype=torch.bfloat16)
        timed("%s %dx%dx%d residual epilogue + rmsnorm" % (tag, M, N, Kd), lambda: K.rmsnorm(K.gemm_residual(x, w, None, res), nw, 1e-6), 2.0 * M * N * Kd, "TFLOP/s")
        timed("  unfused: gemm + rmsnorm(+residual)", lambda: K.rmsnorm(K.gemm(x, w), nw, 1e-6, res=res), 2.0 * M * N * Kd, "TFLOP/s")
    wgu_e = (torch.randn(4, 5632, 1024, device=dev) * 0.02).to(torch.bfloat16)
    timed("grouped gemm_swiglu 4x[~1024,5632,1024]", lambda: K.grouped_gemm_swiglu(xp, wgu_e, offs, 4608, True), 2.0 * 4096 * 5632 * 1024, "TFLOP/s")
    timed("  unfused: grouped gemm + silu_mul", lambda: K.silu_mul(K.grouped_gemm(xp, wgu_e, h1, offs, 0)), 2.0 * 4096 * 5632 * 1024, "TFLOP/s")
    # attention forward: teacher (32 heads, hd 128), student (16 heads, hd 64), CLIP (non-causal 577)
    for (B, Tt, nh, hd, causal, tag) in [(1, T, 32, 128, True, "teacher"), (1, T, 16, 64, True, "student"), (1, 577, 16, 64, False, "CLIP")]:
        qkv = torch.randn(B * Tt, 3 * nh * hd, device=dev).to(torch.bfloat16)
        fl = 4.0 * B * nh * Tt * Tt * hd * (0.5 if causal else 1.0)
        timed("attn fwd %s T%d nh%d hd%d" % (tag, Tt, nh, hd), lambda: K.attention_fwd(qkv, B, Tt, nh, nh, hd, causal), fl, "TFLOP/s")
    # fused KL + CE fwd/bwd, all rows active and the 40%-masked workload
    s = (torch.randn(T, V, device=dev) * 2).to(torch.bfloat16)
    t = (torch.randn(T, V, device=dev) * 2).to(torch.bfloat16)
    for frac, tag in ((0.0, "all rows active"), (0.4, "40% masked (bench workload)")):
        labels = torch.randint(0, V, (T,), device=dev)
        labels[: int(frac * T)] = -100
        active = int(((labels != -100) | torch.cat([labels[1:] != -100, torch.zeros(1, dtype=torch.bool, device=dev)])).sum())
        by = active * 6 * V + (T - active) * 2 * V
        d = torch.empty_like(s)
        timed("kl_fused fwd+bwd [2048,151936] " + tag, lambda: K.kl_fused(s, t, labels, T, V, 1.0, 1.0, False, dlogits=d), by, "GB/s")
    # HBM-bound elementwise kernels at the shapes of the step (teacher H 4096; expert rows 4608 x I 2816)
    xh = torch.randn(T, 4096, device=dev).to(torch.bfloat16); rs = torch.randn_like(xh); wn = torch.ones(4096, device=dev, dtype=torch.bfloat16)
    timed_cold("rmsnorm fwd (+residual) [2048,4096]", lambda: K.RMSNormFn.apply(xh, rs, wn, 1e-6), 4 * T * 4096 * 2, "GB/s")
    qkv7 = torch.randn(T, 3 * 4096, device=dev).to(torch.bfloat16)
    cos = torch.randn(T, 128, device=dev).to(torch.bfloat16); sin = torch.randn(T, 128, device=dev).to(torch.bfloat16)
    pos = torch.arange(T, device=dev)
    timed_cold("rope q|k in place [2048, 2x32x128]", lambda: K.call("lmod_rope", K.ptr(qkv7), qkv7.stride(0), 32, K.ptr(qkv7[:, 4096:]), qkv7.stride(0), 32, 128,
                                                              K.ptr(cos), K.ptr(sin), K.ptr(pos), T, 0), 2 * T * 8192 * 2, "GB/s")
    gu = torch.randn(4608, 5632, device=dev).to(torch.bfloat16); dact = torch.randn(4608, 2816, device=dev).to(torch.bfloat16)
    timed_cold("silu_mul fwd [4608, 2x2816]", lambda: K.silu_mul(gu), 3 * 4608 * 2816 * 2, "GB/s")
    timed_cold("silu_mul bwd [4608, 2x2816]", lambda: K.silu_mul_bwd(dact, gu), 5 * 4608 * 2816 * 2, "GB/s")
    # router + scatter
    x = torch.randn(T, 1024, device=dev).to(torch.bfloat16)
    wg = torch.randn(4, 1024, device=dev) * 0.1
    noise = torch.randn(T, 4, device=dev)
    timed("moe_route_scatter S2048 H1024 E4 (12288 B/token: route+scatter+combine share)", lambda: K.moe_route_scatter(x, wg, noise, 1.5, 0), T * 3 * 1024 * 2, "GB/s")
    r = K.moe_route_scatter(x, wg, noise, 1.5, 0)
    y = torch.randn(r["xp"].shape, device=dev).to(torch.bfloat16)
    timed("moe_gather_combine S2048 H1024", lambda: K.moe_gather_combine(y, r["row"], r["w"], x), T * 4 * 1024 * 2, "GB/s")
    # attention backward
    for (B, Tt, nh, hd) in [(1, T, 16, 64), (1, T, 16, 128)]:
        qkv = torch.randn(B * Tt, 3 * nh * hd, device=dev).to(torch.bfloat16)
        out, lse = K.attention_fwd(qkv, B, Tt, nh, nh, hd, True, need_lse=True)
        go = torch.randn_like(out)
        timed("attn bwd T%d nh%d hd%d" % (Tt, nh, hd), lambda: K.attention_bwd(qkv, out, go, lse, B, Tt, nh, nh, hd, True, hd ** -0.5),
              2.5 * 4.0 * B * nh * Tt * Tt * hd * 0.5, "TFLOP/s")


if __name__ == "__main__":
    main()
