"""Where does a key block's time go inside attn_fwd_kernel?  Runs lmod_attn_fwd_trace (the forward kernel with clock64 stamps at its
pipeline hand-offs) on the teacher / student shapes and prints, per CTA of head 0, the average clocks per key block of every phase of
the softmax warp and of the MMA thread, plus the two hand-off latencies (S ready -> softmax sees it, P arrive -> MMA sees it).
  python profiles/attn_trace.py            (needs a B200; stamps cost a few % -- read the SHARES)"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "llava-mod_b200"))
from llavamod import _C
from llavamod._C import call, ptr

def run(T, nh, nkv, hd, causal=True):
    dev = "cuda"
    qkv = (torch.randn(T, (nh + 2 * nkv) * hd, device=dev) * 0.5).to(torch.bfloat16)
    out = torch.empty(T, nh * hd, device=dev, dtype=torch.bfloat16)
    nqb = (T + 127) // 128
    tr = torch.zeros(nqb, 64, 16, dtype=torch.int64, device=dev)
    for _ in range(3):
        tr.zero_()
        call("lmod_attn_fwd_trace", ptr(qkv), qkv.stride(0), 1, T, nh, nkv, hd, int(causal), hd ** -0.5, ptr(out), out.stride(0), None, ptr(tr))
    torch.cuda.synchronize()
    t = tr.cpu().double()
    print("== T%d nh%d hd%d causal=%d: per key block, clocks (softmax warp 2 | MMA thread | hand-offs)" % (T, nh, hd, causal))
    print("%4s %4s | %7s %7s %7s %7s %7s %7s %7s | %7s %7s %7s %7s | %7s %7s | %7s" % (
        "cta", "nblk", "waitS", "ldS", "maxbar", "exp", "pvwait", "stP", "TOTAL", "qk+kw", "waitP", "waitV", "pv", "S->sm", "P->mma", "mmaTOT"))
    for c in range(nqb):
        x = t[c]
        nblk = int((x[:, 6] > 0).sum())
        if nblk < 4: continue
        x = x[:nblk]
        d = lambda a, b: float((x[1:, a] - x[1:, b]).mean())          # skip block 0 (pipeline fill)
        tot = float((x[1:, 0] - x[:-1, 0]).mean())
        mtot = float((x[1:, 7] - x[:-1, 7]).mean())
        s_to_sm = float((x[1:, 1] - x[:-1, 8]).mean())                # S_j issued at slot 8 of iteration j-1 -> seen by softmax at slot 1 of j
        p_to_mma = float((x[:, 9] - x[:, 6]).clamp(min=-1e6).mean())  # p_full arrive -> MMA thread past its wait
        print("%4d %4d | %7.0f %7.0f %7.0f %7.0f %7.0f %7.0f %7.0f | %7.0f %7.0f %7.0f %7.0f | %7.0f %7.0f | %7.0f" % (
            c, nblk, d(1, 0), d(2, 1), d(3, 2), d(4, 3), d(5, 4), d(6, 5), tot, d(8, 7), d(9, 8), d(10, 9), d(11, 10), s_to_sm, p_to_mma, mtot))

if __name__ == "__main__":
    run(2048, 32, 32, 128)
    run(2048, 16, 16, 64)
    run(577, 16, 16, 64, causal=False)
