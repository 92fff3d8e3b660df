#!/bin/bash
# round-2 GPU session G: full suite (new goldens: airfoil B=2, polygon mesh, integrators; plan-build / pool tests), default bench with all sub-records
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2g_pytest.log
( time timeout 1200 python bench.py --steps 20 --warmup 5 ) 2>gpurun_out/r2g_bench.err | tee gpurun_out/r2g_bench.json | cut -c1-300; tail -4 gpurun_out/r2g_bench.err
