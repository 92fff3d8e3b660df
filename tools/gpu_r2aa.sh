#!/bin/bash
# round-2 GPU session AA: full suite + smoke + the default bench line with every sub-record (as the driver runs it)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2aa_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2aa_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2aa_bench_default_4m.json 2>gpurun_out/r2aa_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2aa_bench_default_4m.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')})
print('e2e',d['e2e']); print('roofline',d['roofline']); print('step_roofline',d.get('step_roofline'))
for k in ('precision_modes','nets','size_sweep','example_meshes','loader_regime','grad_rec_speed','cpu_baseline'):
    print(k, json.dumps(d.get(k))[:900])
PY
