"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of the step).
usage: python profiles/summarize_launches.py gpurun_out/launches_r1.csv [top]"""
import collections
import csv
import re
import sys


def classify(name):
    own = ("kl_", "logp_", "moe_", "rmsnorm", "layernorm", "rope_kernel", "silu_mul", "bias_act", "gelu_bwd", "add_kernel", "splice_",
           "sumsq", "adamw", "softmax_rows", "align_dense", "lmod_gemm", "lmod_attn", "gemm_tcgen05", "gemm2_tcgen05", "attn_fwd", "attn_bwd",
           "attn_dsum", "attn_dq", "rope_vec", "active_rows", "gather_rows", "scatter_rows", "embed_grad", "rmsnorm_wgrad")
    if any(o in name for o in own):
        return "ours"
    if "flash" in name.lower() or "fmha" in name.lower():
        return "lib:attention"
    if any(t in name.lower() for t in ("gemm", "cutlass", "nvjet", "cublas", "xmma", "sm90_", "sm100_")):
        return "lib:gemm"
    return "torch:other"


def main(path, top=40):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        short = re.sub(r"\(.*", "", row["Kernel Name"])[:90]
        tot[short] += v
        cnt[short] += 1
    T = sum(tot.values())
    print("total %.1f us over %d launches (cold-cache, serialised: compare SHARES)" % (T, sum(cnt.values())))
    cls = collections.defaultdict(float)
    for k, v in tot.items():
        cls[classify(k)] += v
    for k, v in sorted(cls.items(), key=lambda x: -x[1]):
        print("  class %-14s %10.1f us %5.1f%%" % (k, v, 100 * v / T))
    for k, v in sorted(tot.items(), key=lambda x: -x[1])[:top]:
        print("%10.1f us %5.1f%%  n=%4d  [%s] %s" % (v, 100 * v / T, cnt[k], classify(k), k))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
