"""Summarise ncu outputs into the small text/JSON files committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r1c.csv > profiles/r1_launches.txt
    python profiles/summarize.py kernel   gpurun_out/k4_r1c.ncu-rep   > profiles/r1_k4_sweep_ncu.txt
"""
import collections
import csv
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_active.avg")


def launches(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    h = rows[0]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        n = r[ki].split("(")[0][:64]
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-66s n=%3d total=%10.3f ms avg=%9.3f ms share=%5.1f%%" % (n, a[0], a[1] / 1e6, a[1] / a[0] / 1e6, 100 * a[1] / tot))


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units, v = rows[0], rows[1], rows[2]
    print("# ncu --set full --clock-control none; kernel:", v[h.index("Kernel Name")][:100])
    for i, n in enumerate(h):
        if n in KEEP or "issue_stalled" in n and "per_issue_active" in n:
            print("%-95s %s %s" % (n, v[i], units[i]))


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
