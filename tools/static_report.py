"""Static evidence about the in-tree libcia.so, produced WITHOUT a GPU (cuobjdump only): registers / stack
(spills) / static shared memory of every kernel, and per-kernel counts of the SASS mnemonics that prove which
hardware paths the code uses (UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA,
UTCBAR = tcgen05.commit, SYNCS = mbarrier, DMMA = fp64 mma.sync, REDUX / MATCH = warp aggregation).

    python tools/static_report.py --ptxas > profiles/r2_static_report.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cell-image-analysis_b200", "libcia.so")
MNEMONICS = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "DMMA",
             "HMMA", "REDUX", "MATCH", "LDGSTS", "FFMA2", "DFMA", "MUFU")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), stdout=subprocess.PIPE, text=True).stdout.split("\n")
    short = []
    for d in out[:len(names)]:
        d = re.sub(r"\(anonymous namespace\)::", "", d)
        d = re.sub(r"^void ", "", d)
        i = d.find("(")
        # keep template arguments, drop the parameter list
        depth, cut = 0, len(d)
        for k, ch in enumerate(d):
            if ch == "<":
                depth += 1
            elif ch == ">":
                depth -= 1
            elif ch == "(" and depth == 0:
                cut = k
                break
        short.append(d[:cut] if i >= 0 else d)
    return short


def main():
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], stdout=subprocess.PIPE, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in res.split("\n"):
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and cur:
            kernels[cur] = dict(reg=int(m.group(1)), stack=int(m.group(2)), shared=int(m.group(3)), local=int(m.group(4)))
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True).stdout
    counts = collections.defaultdict(collections.Counter)
    total = collections.Counter()
    cur = None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["_instr"] += 1
            for mn in MNEMONICS:
                if op == mn or op.startswith(mn):
                    counts[cur][mn] += 1
                    total[mn] += 1
                    break
    names = list(kernels)
    short = demangle(names)
    print(f"# static report of {os.path.relpath(LIB, ROOT)} (cuobjdump --dump-resource-usage / -sass; no GPU involved)")
    print(f"# {len(names)} kernels; library-wide SASS counts: " + ", ".join(f"{k} {v}" for k, v in sorted(total.items(), key=lambda kv: -kv[1])))
    frames = [s for n, s in zip(names, short) if kernels[n]["stack"] or kernels[n]["local"]]
    print(f"# kernels with a stack frame (local arrays or spills, see the ptxas section at the end): {len(frames)}")
    print(f"{'kernel':78s} {'reg':>4s} {'stack':>6s} {'smem':>7s} {'instr':>7s}  hardware-path mnemonics")
    for n, s in sorted(zip(names, short), key=lambda t: t[1]):
        k, c = kernels[n], counts.get(n, {})
        mn = " ".join(f"{m}:{c[m]}" for m in MNEMONICS if c.get(m))
        print(f"{s[:78]:78s} {k['reg']:4d} {k['stack']:6d} {k['shared']:7d} {c.get('_instr', 0):7d}  {mn}")


def ptxas_spills():
    """`nvcc -Xptxas=-v` of every .cu into a scratch directory (the in-tree library is not touched): the kernels
    whose stack frame holds register spills, with the byte counts ptxas reports."""
    import tempfile
    csrc = os.path.join(ROOT, "cell-image-analysis_b200", "csrc")
    tmp = tempfile.mkdtemp(prefix="cia_ptxas_")
    procs = []
    for f in sorted(os.listdir(csrc)):
        if f.endswith(".cu"):
            cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler",
                   "-fPIC", "-Xptxas=-v", "-c", os.path.join(csrc, f), "-o", os.path.join(tmp, f + ".o")]
            procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    rows, clean = [], 0
    for p in procs:
        out = p.communicate()[0]
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", out):
            if int(m.group(3)) or int(m.group(4)):
                rows.append((m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4))))
            else:
                clean += 1
    print()
    print(f"# ptxas -v (nvcc {' '.join(['-O3', '-lineinfo', 'sm_100a'])}): {clean} kernels without spills, {len(rows)} with:")
    for (n, st, a, b), s in sorted(zip(rows, demangle([r[0] for r in rows])), key=lambda t: t[1]):
        print(f"{s[:78]:78s} stack {st:4d} B, spill stores {a:4d} B, spill loads {b:4d} B")


if __name__ == "__main__":
    main()
    if "--ptxas" in sys.argv:
        ptxas_spills()
