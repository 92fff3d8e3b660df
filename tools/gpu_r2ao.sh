#!/bin/bash
# round-2 GPU session AO: ncu --set full of the fused edge-MLP backward call with the node-level layer 1 (traffic of its 4 kernels)
mkdir -p gpurun_out
timeout 300 python tools/edge_bwd_profile.py 4000000 f16
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mlp_tc_bwd|dz_incidence" -s 8 -c 4 -f -o gpurun_out/r2ao_prof_edge_bwd \
  python tools/edge_bwd_profile.py 4000000 f16 > gpurun_out/r2ao_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r2ao_ncu.log
python tools/ncu_summary.py gpurun_out/r2ao_prof_edge_bwd.ncu-rep "ncu --set full --clock-control none, python tools/edge_bwd_profile.py 4000000 f16 (the bench's 4 M-cell quad mesh: 8 M edges, 4 M nodes), the four kernels of one fused edge-MLP backward call (node-level layer 1, 16-bit gradient streams); round 2" > gpurun_out/r2ao_ncu_edge_bwd_grad16_8Medges.csv
cut -d, -f1-4 gpurun_out/r2ao_ncu_edge_bwd_grad16_8Medges.csv
