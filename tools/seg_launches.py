"""ncu launch list (csv of `--metrics gpu__time_duration.sum`) of tests/diagnostics/seg_once.py -> the committed text summary:
  python tools/seg_launches.py gpurun_out/seg_launches_<tag>.csv > profiles/<tag>_seg_launches.txt
Prints the LAST of the passes the script made (one 2048 x 2048 field: normalize, U-Net, instances)."""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, gi, bi = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Grid Size", "Block Size"))
out = [(r[ki], float(r[vi].replace(",", "")), r[gi], r[bi]) for r in rows[1:]]
first = [i for i, o in enumerate(out) if "seg_hist_kernel" in o[0]]
out = out[first[-1]:] if first else out
tot = sum(v for _, v, _, _ in out)
print("# ncu --metrics gpu__time_duration.sum --clock-control none, python tests/diagnostics/seg_once.py (last pass):")
print("# one 2048 x 2048 field: normalize, U-Net (2D_versatile_fluo topology), instances (58 050 candidates -> 529 labels)")
groups = {"normalize": 0.0, "unet": 0.0, "instances": 0.0}
for name, v, g, b in out:
    short = name.replace("<unnamed>::", "").replace("void ", "")
    key = "normalize" if any(k in short for k in ("seg_hist", "seg_percentile", "seg_normalize")) else \
          "unet" if any(k in short for k in ("seg_first", "seg_conv")) else "instances"
    groups[key] += v
    print(f"{v / 1000:10.1f} us {100 * v / tot:5.1f} %  grid {g:>14s} block {b:>12s}  {short[:100]}")
print(f"{tot / 1000:10.1f} us total  (" + ", ".join(f"{k} {v / 1000:.1f} us" for k, v in groups.items()) + ")")
