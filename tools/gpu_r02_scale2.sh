#!/bin/bash
# round 2, scaling call of the final build: config 5 (default flags, with the end-to-end leg) and config 4 on $1 GPUs
mkdir -p gpurun_out
n=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2958$n \
  bench.py --gpus $n --no-cpu > gpurun_out/t_bench_c5_${n}gpu.json 2> gpurun_out/t_bench_c5_${n}gpu.err
echo "c5 $n rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2959$n \
  bench.py --gpus $n --config c4 --no-cpu > gpurun_out/t_bench_c4_${n}gpu.json 2> gpurun_out/t_bench_c4_${n}gpu.err
echo "c4 $n rc=$?"
python - $n <<'PY'
import json,sys
n=int(sys.argv[1])
for c in ('c5','c4'):
    try:
        d=json.loads(open(f'gpurun_out/t_bench_{c}_{n}gpu.json').read().strip().splitlines()[-1])
        print(c,n,'ms_per_step %.2f'%d['ms_per_step'],'value %.3e'%d['value'],'e2e',(d.get('e2e') or {}).get('seconds_per_step'), d['rank0_phases_ms'])
    except Exception as ex: print(c,n,'failed',ex)
PY
tail -3 gpurun_out/t_bench_c5_${n}gpu.err
