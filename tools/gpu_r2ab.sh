#!/bin/bash
# round-2 GPU session AB (N GPUs): data-parallel bench line with the `cells` strong-scaling sub-record; partition parity check
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2ab_bench_dp${N}_and_cells${N}_4m.json 2>gpurun_out/r2ab_bench${N}.err; echo "bench rc=$?"
python - $N <<'PY'
import json,sys
n=sys.argv[1]
d=json.loads(open(f'gpurun_out/r2ab_bench_dp{n}_and_cells{n}_4m.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','scaling','gpu_launches','clocks')}); print('e2e',d.get('e2e'))
print('cells',json.dumps(d.get('cells'))[:1500])
PY
