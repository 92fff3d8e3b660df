#!/bin/bash
# round-2 GPU session A: regression tests, parity table of all modes, kernel timings, 4 M-cell bench in bf16 and f16
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2a_pytest.log
timeout 900 python tools/parity_report.py gpurun_out/r2a_parity_report.json > gpurun_out/r2a_parity.log 2>&1; echo "parity rc=$?"
for p in bf16 f16; do for m in EDGE NODE; do timeout 120 python tools/tc_profile.py 8000000 $m $p; done; done 2>&1 | tee gpurun_out/r2a_tc_profile.log
timeout 300 python tools/reduce_bench.py 2000 2>&1 | tee gpurun_out/r2a_reduce_bench.log
for p in bf16 f16; do timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --precision $p 2>gpurun_out/r2a_bench_$p.err | tee gpurun_out/r2a_bench_$p.json | cut -c1-400; done
