#!/bin/bash
# round 2, call D (1 GPU): batched Richardson-Lucy -- deconvolution / slab / chain tests and the config-5 bench
mkdir -p gpurun_out
python -m pytest tests/test_deconv_gpu.py tests/test_slab_gpu.py tests/test_chain_driver_gpu.py -m gpu -q -s -x > gpurun_out/d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/d_pytest.log
grep -E "passed|failed|rel err|RL band|config 3|FAILED|Error|error" gpurun_out/d_pytest.log | tail -20
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/d_bench_c5.json 2> gpurun_out/d_bench_c5.err
echo "bench rc=$?" >> gpurun_out/d_bench_c5.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/d_bench_c5.json').read().strip().splitlines()[-1])
print('ms_per_step',d['ms_per_step']); print({k:(v.get('ms')) for k,v in d['stage_breakdown'].items() if isinstance(v,dict) and 'ms' in v}); print(d['stage_breakdown'].get('stage_totals_ms')); print(d['gpu_launches'])
PY
tail -3 gpurun_out/d_bench_c5.err
