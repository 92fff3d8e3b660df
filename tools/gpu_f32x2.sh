#!/bin/bash
# A/B of the packed-f32x2 epilogues (libfvgn_b200.so) against the scalar build (libfvgn_b200_scalar.so, -DFVGN_F32X2=0):
# tcgen05 kernel tests + parity with the packed build, then kernel timings and a 1 M-cell bench with both.
tag=${1:-x}
mkdir -p gpurun_out
L=gen_fvgn_steady_b200
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_f32x2_$tag.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/pytest_f32x2_$tag.log
run() {  # $1 = label
  for m in EDGE NODE; do timeout 120 python tools/tc_profile.py 2000000 $m bf16; done
  timeout 300 python bench.py --cells 1000000 --steps 10 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', 'bench 1M cells: ms/step', round(d['ms_per_step'],3), 'clocks', d['clocks'])"
}
echo "== packed"; run packed 2>&1 | tee gpurun_out/f32x2_packed_$tag.log
cp $L/libfvgn_b200.so $L/libfvgn_b200_packed.so; cp $L/libfvgn_b200_scalar.so $L/libfvgn_b200.so
echo "== scalar"; run scalar 2>&1 | tee gpurun_out/f32x2_scalar_$tag.log
cp $L/libfvgn_b200_packed.so $L/libfvgn_b200.so
echo "== packed again"; run packed2 2>&1 | tee gpurun_out/f32x2_packed2_$tag.log
