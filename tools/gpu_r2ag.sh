#!/bin/bash
# round-2 GPU session AG (final profiles): default bench line; ncu launch list of the same command; kernel table of TransFVGN_v2
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2ag_bench_default_4m.json 2>gpurun_out/r2ag_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2ag_bench_default_4m.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')})
print('e2e',d['e2e']); print('roofline',{k:d['roofline'][k] for k in ('achieved','ms_per_launch','frac')}); print('step_roofline',d.get('step_roofline'))
for k in ('precision_modes','nets','size_sweep','loader_regime','cpu_baseline'):
    print(k, json.dumps(d.get(k))[:700])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2000 --csv --log-file gpurun_out/r2ag_launches_f16_4m.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2ag_ncu_bench.log 2>&1; echo "ncu launch list rc=$?"
python tools/launch_summary.py gpurun_out/r2ag_launches_f16_4m.csv 0 | tee gpurun_out/r2ag_launches_f16_4m_summary.txt | head -24
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --net TransFVGN_v2 --mp 3 --kernel-summary gpurun_out/r2ag_kernels_v2_4m.txt 2>/dev/null | cut -c1-200
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | cut -c1-600
