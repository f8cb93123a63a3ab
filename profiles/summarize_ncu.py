"""One line per profiled launch of an `ncu --set full` report: the metrics DESIGN.md quotes.
    python profiles/summarize_ncu.py gpurun_out/kernels_r2.ncu-rep > profiles/kernels_r2_ncu_summary.txt"""
import csv
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "us", 1.0),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%", 1.0),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%", 1.0),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1.0),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 1.0),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1.0),
        ("dram__bytes_read.sum", "rdMB", None), ("dram__bytes_write.sum", "wrMB", None),
        ("lts__t_sector_hit_rate.pct", "L2hit%", 1.0),
        ("launch__registers_per_thread", "regs", 1.0), ("launch__grid_size", "grid", 1.0), ("smsp__inst_executed.sum", "Minst", 1e-6)]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    name_i = h.index("Kernel Name")
    print("# ncu --set full --clock-control none, %s (per launch; cold-cache, serialised: compare shares, not absolutes)" % path)
    print("%-44s " % "kernel" + " ".join("%8s" % c[1] for c in COLS))
    for r in rows[2:]:
        vals = []
        for key, label, scale in COLS:
            if key not in h:
                vals.append("%8s" % "-")
                continue
            i = h.index(key)
            try:
                v = float(r[i])
            except ValueError:
                vals.append("%8s" % "-")
                continue
            if scale is None:                      # bytes with a unit column
                u = units[i].lower()
                v = v * {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, 1e-6)
            else:
                v *= scale
            vals.append("%8.1f" % v)
        nm = r[name_i].replace("void ", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
        print("%-44s " % nm[:44] + " ".join(vals))


if __name__ == "__main__":
    main(sys.argv[1])
