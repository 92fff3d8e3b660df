#!/bin/bash
# round-2 GPU session Z: dz_incidence_kernel variant sweep (library built with -DFVGN_DZI_SWEEP)
mkdir -p gpurun_out
for v in 0 1 2 3 4 5; do FVGN_DZI_VARIANT=$v timeout 300 python tools/dzi_sweep.py 2>/dev/null | grep dz_incidence; done | tee gpurun_out/r2z_dzi_sweep.txt
