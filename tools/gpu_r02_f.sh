#!/bin/bash
# round 2, call F (8 GPUs): scaling point at N = 8 (and 4) with the slab Richardson-Lucy
mkdir -p gpurun_out
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n \
  bench.py --gpus $n --steps 5 --warmup 3 --no-cpu > gpurun_out/f_bench_${n}gpu.json 2> gpurun_out/f_bench_${n}gpu.err
echo "bench rc=$?" >> gpurun_out/f_bench_${n}gpu.err
python - $n <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/f_bench_{n}gpu.json').read().strip().splitlines()[-1])
    print(n,'GPUs ms_per_step',d['ms_per_step'],'value',d['value']); print(d['rank0_phases_ms']); print({k:(v.get('ms')) for k,v in d['stage_breakdown'].items() if isinstance(v,dict) and 'ms' in v}); print('e2e',d['e2e'])
except Exception as ex: print('no bench line', ex)
PY
tail -4 gpurun_out/f_bench_${n}gpu.err
done
