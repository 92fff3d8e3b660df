#!/bin/bash
# Round-end check of the committed state: full GPU suite, smoke, headline bench (with CPU baseline), TransFVGN_v2 bench, launch list.
tag=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "gpu suite rc=$?"; tail -2 gpurun_out/pytest_gpu_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$tag.log
timeout 600 python bench.py > gpurun_out/bench_4m_$tag.log 2>&1; echo "bench rc=$?"; grep '^{' gpurun_out/bench_4m_$tag.log | cut -c1-260
timeout 300 python bench.py --net TransFVGN_v2 --mp 3 --cells 1000000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v2_1m_$tag.log 2>&1; echo "bench v2 rc=$?"
grep '^{' gpurun_out/bench_v2_1m_$tag.log | cut -c1-260
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_epd_1m_$tag.csv \
  python bench.py --cells 1000000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_epd_$tag.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/launches_epd_1m_$tag.csv | tee gpurun_out/launches_epd_1m_${tag}_summary.txt | head -24
