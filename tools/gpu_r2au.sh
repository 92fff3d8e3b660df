#!/bin/bash
# round-2 GPU session AU: the default bench line once more on the final code (roofline call with the product's streams, warm-up floor)
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2au_bench_default_4m.json 2>gpurun_out/r2au_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2au_bench_default_4m.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')})
print('e2e',d['e2e']); print('roofline',{k:d['roofline'][k] for k in ('kernel','achieved','ms_per_launch','frac','traffic')}); print('step_roofline',d.get('step_roofline'))
for k in ('precision_modes','nets','loader_regime'):
    print(k, json.dumps(d.get(k))[:600])
print('example', {k:(round(v['eager']['ms_per_step'],2), round(v['cuda_graph']['ms_per_step'],2)) for k,v in d['example_meshes'].items()})
PY
