#!/bin/bash
# round-2 GPU session AK: ncu --set full with source correlation of the forward EDGE kernel (16-bit latent streams)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_fwd -s 4 -c 1 -f -o gpurun_out/r2ak_prof_fwd_edge \
  python tools/tc_profile.py 8000000 EDGE f16 > gpurun_out/r2ak_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r2ak_ncu.log
ls -la gpurun_out/r2ak_prof_fwd_edge.ncu-rep
