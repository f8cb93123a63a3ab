"""KL kernel alone: LMOD_KL_MODE={sb256,db256,db512} python profiles/kl_bench.py  (HBM-bound: algorithmic bytes / CUDA-event time)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "llava-mod_b200"))
from llavamod import kernels as K
T, V = 2048, 151936
dev = "cuda"
s = (torch.randn(T, V, device=dev) * 2).to(torch.bfloat16)
t = (torch.randn(T, V, device=dev) * 2).to(torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for frac, tag in ((0.0, "all rows active"), (0.4, "40% masked head"), (0.57, "bench-like: 57% masked head")):
    labels = torch.randint(0, V, (T,), device=dev)
    labels[: int(frac * T)] = -100
    active = int(((labels != -100) | torch.cat([labels[1:] != -100, torch.zeros(1, dtype=torch.bool, device=dev)])).sum())
    by = active * 6 * V + (T - active) * 2 * V
    compact = os.environ.get("COMPACT", "1") == "1"
    if compact:                      # what the trainer does: the kernel only sees the supervised rows
        rows = K.active_rows(labels, T)
        s_in, t_in = K.gather_rows(s, *rows), K.gather_rows(t, *rows)
        by = active * 6 * V
    else:
        rows, s_in, t_in = None, s, t
    d = torch.empty_like(s_in)
    for _ in range(3):
        K.kl_fused(s_in, t_in, labels, T, V, 1.0, 1.0, False, dlogits=d, rows=rows)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); K.kl_fused(s_in, t_in, labels, T, V, 1.0, 1.0, False, dlogits=d, rows=rows); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    print("%-8s %-7s %-32s %.3f ms  %.0f GB/s (whole kl_fused call: counts + fused + finalize)" % (os.environ.get("LMOD_KL_MODE", "default"), "compact" if compact else "dense", tag, ms, by / ms / 1e6), flush=True)
