#!/bin/bash
# round 2, call O (2 GPUs): multi-GPU paths with the fused trace + energy kernel and the spectral hand-off
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_slab_gpu.py -m gpu -q > gpurun_out/o_pytest_slab.log 2>&1
echo "pytest rc=$?" >> gpurun_out/o_pytest_slab.log
tail -5 gpurun_out/o_pytest_slab.log
n=2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29582 \
  bench.py --gpus $n --no-cpu > gpurun_out/o_bench_c5_${n}gpu.json 2> gpurun_out/o_bench_c5_${n}gpu.err
echo "c5 $n rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29592 \
  bench.py --gpus $n --config c4 --no-cpu > gpurun_out/o_bench_c4_${n}gpu.json 2> gpurun_out/o_bench_c4_${n}gpu.err
echo "c4 $n rc=$?"
python - <<'PY'
import json
for c in ('c5','c4'):
    for n in (2,):
        try:
            d=json.loads(open(f'gpurun_out/o_bench_{c}_{n}gpu.json').read().strip().splitlines()[-1])
            print(c,n,'ms_per_step %.2f'%d['ms_per_step'],'value %.3e'%d['value'],'e2e',(d.get('e2e') or {}).get('seconds_per_step'), d['rank0_phases_ms'])
            print({k:round(v.get('ms'),2) for k,v in d['stage_breakdown'].items() if isinstance(v,dict) and 'ms' in v})
        except Exception as ex: print(c,n,'failed',ex)
PY
tail -3 gpurun_out/o_bench_c5_2gpu.err
