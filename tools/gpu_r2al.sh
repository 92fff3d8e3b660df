#!/bin/bash
# round-2 GPU session AL: final check of the driver's sequence: GPU tests, smoke(), default bench line
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2al_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2al_pytest.log | cut -c1-200
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -4
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2al_bench_default_4m.json 2>gpurun_out/r2al_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2al_bench_default_4m.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(d['gpu_launches_detail'])
print('e2e',d['e2e']); print('step_roofline',d.get('step_roofline'))
for k in ('nets','size_sweep'):
    print(k, json.dumps(d.get(k))[:500])
PY
