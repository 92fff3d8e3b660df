#!/bin/bash
# round-2 GPU session AR: ncu --set full of backward kernel B (NODE and EDGE) with fp32 and with 16-bit gradient streams
mkdir -p gpurun_out
for cfg in "0 gs0" "16 gs1"; do set -- $cfg
timeout 300 ncu --set full --clock-control none -k regex:mlp_tc_bwd_b_kernel -s $1 -c 2 -f -o gpurun_out/r2ar_prof_b_$2 \
  python tools/grad16_bench.py 4000000 > gpurun_out/r2ar_ncu_$2.log 2>&1; echo "ncu $2 rc=$?"
python tools/ncu_summary.py gpurun_out/r2ar_prof_b_$2.ncu-rep "ncu --set full, backward kernel B, $2, 4 M-cell mesh" > gpurun_out/r2ar_ncu_b_$2.csv
done
cat gpurun_out/r2ar_ncu_b_gs0.csv gpurun_out/r2ar_ncu_b_gs1.csv | grep -v "^#" | cut -c1-330
