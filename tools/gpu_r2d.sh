#!/bin/bash
# round-2 GPU session D: all-mode parity tests, smoke in the benchmarked modes, default bench with sub-records
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2d_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
( time timeout 900 python bench.py --steps 20 --warmup 5 ) 2>gpurun_out/r2d_bench.err | tee gpurun_out/r2d_bench.json | cut -c1-400; tail -5 gpurun_out/r2d_bench.err
