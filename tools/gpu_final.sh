#!/bin/bash
# Round-end GPU call: full GPU suite, smoke, headline bench, TransFVGN_v2 benches, launch list, ncu capture of the Transolver kernels.
tag=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "gpu suite rc=$?"; tail -2 gpurun_out/pytest_gpu_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$tag.log
timeout 600 python bench.py > gpurun_out/bench_4m_$tag.log 2>&1; echo "bench rc=$?"; grep '^{' gpurun_out/bench_4m_$tag.log | cut -c1-260
for c in 1000000 4000000; do
  timeout 300 python bench.py --net TransFVGN_v2 --mp 3 --cells $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v2_${c}_$tag.log 2>&1; echo "bench v2 $c rc=$?"
  grep '^{' gpurun_out/bench_v2_${c}_$tag.log | cut -c1-260
done
timeout 120 python tools/ts_profile.py 1000000 1 5 | tee gpurun_out/ts_times_$tag.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'ts_' -s 18 -c 9 -f -o gpurun_out/prof_ts_$tag \
  python tools/ts_profile.py 1000000 1 1 > gpurun_out/ncu_full_ts_$tag.log 2>&1; echo "ncu full rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_v2_$tag.csv \
  python bench.py --net TransFVGN_v2 --mp 3 --cells 1000000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_v2_$tag.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/launches_v2_$tag.csv | tee gpurun_out/launches_v2_${tag}_summary.txt | head -34
