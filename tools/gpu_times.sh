#!/bin/bash
# per-kernel durations of the fused MLP kernels (ncu, cold-cache) + wall timings.  Usage: tools/gpu_times.sh tag [rows]
tag=${1:-x}; rows=${2:-2000000}
mkdir -p gpurun_out
for m in EDGE NODE; do
  timeout 120 python tools/tc_profile.py $rows $m bf16
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_active --clock-control none -k regex:mlp_tc -s 6 -c 3 --csv --log-file gpurun_out/k_${m}_$tag.csv python tools/tc_profile.py $rows $m bf16 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open("gpurun_out/k_${m}_$tag.csv") if l.startswith('"'))]
h=rows[0]; k=h.index("Kernel Name"); mn=h.index("Metric Name"); v=h.index("Metric Value"); i=h.index("ID")
d={}
for r in rows[1:]:
    d.setdefault((r[i],r[k][:60]),{})[r[mn]]=r[v]
for (id_,kn),m in d.items():
    print(kn, {a.split('.')[0][-22:]:b for a,b in m.items()})
PY
done 2>&1 | tee gpurun_out/times_$tag.log
