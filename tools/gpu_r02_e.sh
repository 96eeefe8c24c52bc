#!/bin/bash
# round 2, call E (1 GPU): tensor-core edge energies -- parity test, then the config-5 bench with THZ_EDGE_MMA=on
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_edges_mma_gpu.py -m gpu -q -s -x > gpurun_out/e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/e_pytest.log
tail -30 gpurun_out/e_pytest.log
THZ_EDGE_MMA=on timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/e_bench_c5.json 2> gpurun_out/e_bench_c5.err
echo "bench rc=$?" >> gpurun_out/e_bench_c5.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/e_bench_c5.json').read().strip().splitlines()[-1])
    print('ms_per_step',d['ms_per_step']); print({k:(v.get('ms')) for k,v in d['stage_breakdown'].items() if isinstance(v,dict) and 'ms' in v})
except Exception as ex: print('no bench line', ex)
PY
tail -3 gpurun_out/e_bench_c5.err
