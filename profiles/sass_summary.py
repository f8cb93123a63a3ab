"""Per-kernel SASS instruction counts of liblmod_b200.so (cuobjdump -sass): the evidence that the hot kernels are Blackwell-native
(UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UBLKCP = TMA, no HMMA = no legacy mma.sync path).
    python profiles/sass_summary.py > profiles/sass_summary_r2.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "llava-mod_b200", "llavamod", "liblmod_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "MUFU", "FFMA2", "FADD2", "FMUL2", "SYNCS", "BAR", "RED", "ATOM"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    fn, rows = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r"\(anonymous namespace\)::", "", fn)
            fn = re.sub(r"\(.*$", "", fn)[:70]
            rows[fn] = collections.Counter()
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            op = m.group(1)
            rows[fn]["total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    rows[fn][k] += 1
                    break
    print("# cuobjdump -sass %s  (instruction counts per kernel; static code, not executed counts)" % os.path.relpath(SO, ROOT))
    print("%-72s %6s  %s" % ("kernel", "total", "  ".join("%s" % k for k in KEYS)))
    tot = collections.Counter()
    for fn, c in rows.items():
        print("%-72s %6d  %s" % (fn, c["total"], "  ".join("%*d" % (len(k), c[k]) for k in KEYS)))
        tot.update(c)
    print("%-72s %6d  %s" % ("ALL KERNELS", tot["total"], "  ".join("%*d" % (len(k), tot[k]) for k in KEYS)))
    assert tot["HMMA"] == 0, "legacy mma.sync code found"


if __name__ == "__main__":
    main()
