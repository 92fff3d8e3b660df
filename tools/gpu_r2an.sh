#!/bin/bash
# round-2 GPU session AN (final, N = 1): GPU tests, smoke(), default bench line, ncu launch list, kernel tables
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2an_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" gpurun_out/r2an_pytest.log | tail -1
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -4
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2an_bench_default_4m.json 2>gpurun_out/r2an_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2an_bench_default_4m.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(d['gpu_launches_detail'])
print('e2e',d['e2e']); print('roofline',{k:d['roofline'][k] for k in ('achieved','ms_per_launch','frac')}); print('step_roofline',d.get('step_roofline'))
for k in ('precision_modes','nets','size_sweep','example_meshes','loader_regime','cpu_baseline'):
    print(k, json.dumps(d.get(k))[:1200])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2000 --csv --log-file gpurun_out/r2an_launches_f16_4m.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2an_ncu_bench.log 2>&1; echo "ncu launch list rc=$?"
python tools/launch_summary.py gpurun_out/r2an_launches_f16_4m.csv 0 | tee gpurun_out/r2an_launches_f16_4m_summary.txt | head -12
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --kernel-summary gpurun_out/r2an_kernels_f16_4m.txt 2>/dev/null | cut -c1-160
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --net TransFVGN_v2 --mp 3 --kernel-summary gpurun_out/r2an_kernels_v2_4m.txt 2>/dev/null | cut -c1-160
timeout 300 python tools/edge_bwd_profile.py 4000000 f16 2>&1 | tail -1
