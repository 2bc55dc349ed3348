"""Turn the ncu outputs brought back in gpurun_out/ into the committed text summaries.

  python profiles/summarize.py launches gpurun_out/launches_<tag>.csv  > profiles/<tag>_launches.txt
  python profiles/summarize.py full     gpurun_out/prof_<tag>.ncu-rep  > profiles/<tag>_full.txt
"""
import collections
import csv
import re
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
    "sm__inst_executed_pipe_tensor.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
]


def short(name):
    m = re.search(r"conv3x3_kernel<([^>]*)>", name)
    if m:
        return "conv3x3<" + m.group(1).replace(" ", "") + ">"
    m = re.search(r"(\w+)<([^>]*)>\(", name)
    if m and "::" in name:
        return m.group(1) + "<" + m.group(2).replace(" ", "") + ">"
    name = re.sub(r"\(.*", "", name)
    return re.sub(r".*::", "", name).replace("void ", "")


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        a = agg.setdefault(short(r[kn]), [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none ({path}); "
          f"cold-cache serialised launches: compare SHARES\n# total {tot / 1e6:.2f} ms over "
          f"{sum(v[0] for v in agg.values())} launches")
    print(f"{'kernel':48s} {'n':>5s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:48s} {v[0]:5d} {v[1] / 1e3:12.1f} {v[1] / 1e3 / v[0]:10.1f} {v[1] / tot * 100:6.1f}%")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none ({path})")
    for r in rows[2:]:
        print(f"\n== {short(r[hdr.index('Kernel Name')])}")
        for m in FULL_METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"  {m:64s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
