"""Summarise ncu outputs into small tracked files under profiles/.

    python tools/summarize_ncu.py launches <launches.csv> <out.md>     # per-kernel share of a launch list
    python tools/summarize_ncu.py full <report.ncu-rep> <out.md>       # key metrics of an `ncu --set full` capture
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct"]


def launches(src, dst):
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr = rows[0]
    i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        n = r[i_name].split("(")[0][:80]
        v = float(r[i_val].replace(",", ""))
        v = v / 1e3 if r[i_unit] == "ns" else (v * 1e3 if r[i_unit] == "ms" else v)
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none): per-kernel share\n\n")
        f.write("source: %s; %d launches, %.1f us total (cold-cache, serialised: compare shares, not absolutes)\n\n" %
                (src, sum(a[0] for a in agg.values()), tot))
        f.write("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f%% | %.2f |\n" % (n, a[0], a[1], 100 * a[1] / tot, a[1] / a[0]))


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write("# ncu --set full capture: key metrics per kernel\n\nsource: %s\n\n" % src)
        for r in rows[2:]:
            f.write("## `%s`\n\n| metric | value | unit |\n|---|---:|---|\n" % r[idx["Kernel Name"]][:100])
            for k in KEYS:
                if k in idx:
                    f.write("| %s | %s | %s |\n" % (k, r[idx[k]], units[idx[k]]))
            f.write("\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
