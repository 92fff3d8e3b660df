#!/bin/bash
# Quick GPU check: full GPU suite + launch list of the 1 M-cell EPD step.  Usage: tools/gpu_quick.sh tag
tag=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "gpu suite rc=$?"; tail -4 gpurun_out/pytest_gpu_$tag.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_epd_1m_$tag.csv \
  python bench.py --cells 1000000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_epd_$tag.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/launches_epd_1m_$tag.csv | tee gpurun_out/launches_epd_1m_${tag}_summary.txt | head -16
