"""Per-source-line instruction counts of one captured launch:
  ncu -i X.ncu-rep --page source --csv --print-source cuda,sass [--launch-skip K --launch-count 1] > src.csv
  python profiles/source_lines.py src.csv [top_n]
Rows with a line number and '-' as address are the per-line aggregates."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(r for r in rows if "Instructions Executed" in r)
ie, ln = hdr.index("Instructions Executed"), hdr.index("Line No")
src, addr = hdr.index("Source"), hdr.index("Address")
smp = hdr.index("# Samples")
out, tot, tots = [], 0, 0
for r in rows:
    if len(r) <= ie or r[addr] != "-" or not r[ln].isdigit():
        continue
    try:
        v = int(r[ie]); s = int(r[smp])
    except ValueError:
        continue
    tot += v; tots += s
    out.append((v, s, int(r[ln]), r[src][:105]))
print(f"total warp instructions {tot}, stall samples {tots}")
for v, s, l, t in sorted(out, reverse=True)[:top]:
    print(f"{v:11d} {100 * v / tot:5.1f}%  samples {100 * s / max(tots, 1):5.1f}%  L{l}: {t}")
