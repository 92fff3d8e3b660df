#!/bin/bash
# One GPU call for the Transolver kernels: block tests, whole GPU suite, TransFVGN_v2 bench at 1 M cells, ncu launch list.
# Usage: tools/gpu_ts.sh [tag] [cells]
tag=${1:-x}; cells=${2:-1000000}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_transolver.py -x -q -m gpu -s > gpurun_out/pytest_ts_$tag.log 2>&1; echo "ts rc=$?"
tail -4 gpurun_out/pytest_ts_$tag.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "gpu suite rc=$?"
tail -4 gpurun_out/pytest_gpu_$tag.log
timeout 300 python bench.py --net TransFVGN_v2 --mp 3 --cells $cells --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v2_$tag.log 2>&1; echo "bench rc=$?"
grep '^{' gpurun_out/bench_v2_$tag.log | cut -c1-330
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_v2_$tag.csv \
  python bench.py --net TransFVGN_v2 --mp 3 --cells $cells --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_v2_$tag.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/launches_v2_$tag.csv | tee gpurun_out/launches_v2_${tag}_summary.txt | head -40
# one full-metrics capture of the slice / de-slice kernels (first step), read back with tools/ncu_lines.py / ncu -i
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'ts_slice|ts_deslice' -c 8 -f -o gpurun_out/prof_ts_$tag \
  python bench.py --net TransFVGN_v2 --mp 3 --cells $cells --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_ts_$tag.log 2>&1; echo "ncu full rc=$?"
