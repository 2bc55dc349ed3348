#!/bin/bash
# DRAM bytes, duration and tensor-pipe activity of every CAE kernel of one bench pass (7565 cells per launch)
# usage (GPU box): bash tools/cae_traffic_ncu.sh <tag>
TAG=$1
CMD="python bench.py --steps 1 --warmup 1 --fields 32 --pool 16 --chunk 16 --no-cpu-baseline --no-svm-extras"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_traffic_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:'conv|final_tapsum' -s 14 -c 14 --csv --log-file gpurun_out/cae_traffic_$TAG.csv $CMD > gpurun_out/ncu_traffic_$TAG.log 2>&1
python - <<P
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/cae_traffic_$TAG.csv", errors="ignore")) if len(r) > 14 and r[0].isdigit()]
k=collections.OrderedDict()
for r in rows:
    k.setdefault((r[0], r[4].split("(")[0].replace("void ","").replace("<unnamed>::","")), {})[r[12]] = (float(r[14]), r[13])
print(f"{'kernel':44s} {'us':>8s} {'rd MB':>8s} {'wr MB':>8s} {'tensor%':>8s} {'tc%':>6s} {'opnd%':>6s} {'issue%':>7s}")
for (i, name), m in k.items():
    g=lambda s: next((v for kk,(v,u) in m.items() if s in kk), float('nan'))
    conv=lambda s: next((v*(1e-3 if u=='ns' else 1) for kk,(v,u) in m.items() if s in kk), float('nan'))
    def mb(s):
        for kk,(v,u) in m.items():
            if s in kk: return v*{'byte':1e-6,'Kbyte':1e-3,'Mbyte':1,'Gbyte':1e3}.get(u,1)
        return float('nan')
    print(f"{name[:44]:44s} {conv('time_duration'):8.1f} {mb('bytes_read'):8.1f} {mb('bytes_write'):8.1f} {g('pipe_tensor_cycles'):8.1f} {g('pipe_tc_cycles'):6.1f} {g('wavefronts_mem_shared'):6.1f} {g('issue_active'):7.1f}")
P
