#!/bin/bash
# round 2, call H (1 GPU): slab / deconvolution / edge tests after the scheduling changes, config-5 bench
mkdir -p gpurun_out
python -m pytest tests/test_deconv_gpu.py tests/test_slab_gpu.py tests/test_edges_mma_gpu.py tests/test_handoff_gpu.py -m gpu -q -x > gpurun_out/h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/h_pytest.log
tail -5 gpurun_out/h_pytest.log
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/h_bench_c5.json 2> gpurun_out/h_bench_c5.err
echo "bench rc=$?" >> gpurun_out/h_bench_c5.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/h_bench_c5.json').read().strip().splitlines()[-1])
print('ms_per_step',d['ms_per_step']); print({k:(v.get('ms')) for k,v in d['stage_breakdown'].items() if isinstance(v,dict) and 'ms' in v})
PY
tail -3 gpurun_out/h_bench_c5.err
