#!/bin/bash
# round-2 GPU session P (profiles): ncu launch list of the default bench command, ncu --set full of the MLP kernels and the reductions
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2000 --csv --log-file gpurun_out/r2p_launches_f16_4m.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2p_ncu_bench.log 2>&1; echo "ncu launch list rc=$?"
python tools/launch_summary.py gpurun_out/r2p_launches_f16_4m.csv 0 | tee gpurun_out/r2p_launches_f16_4m_summary.txt | head -30
for m in EDGE NODE; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_tc -s 6 -c 3 -f -o gpurun_out/r2p_prof_$m \
    python tools/tc_profile.py 8000000 $m f16 > gpurun_out/r2p_ncu_$m.log 2>&1; echo "ncu full $m rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pipe_reduce -s 24 -c 6 -f -o gpurun_out/r2p_prof_reduce \
  python tools/reduce_variants.py run 2000 > gpurun_out/r2p_ncu_reduce.log 2>&1; echo "ncu full reduce rc=$?"
ls -la gpurun_out/r2p_*
