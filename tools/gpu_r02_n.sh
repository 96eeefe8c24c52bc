#!/bin/bash
# round 2, call N (1 GPU): spectral hand-off -- fused-kernel tests, rest of the GPU suite, config-5 bench with and without it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chain_fused_gpu.py -m gpu -q > gpurun_out/n_pytest_fused.log 2>&1
echo "pytest rc=$?" >> gpurun_out/n_pytest_fused.log
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/n_pytest_fused.log | head -40
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_chain_fused_gpu.py > gpurun_out/n_pytest_rest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/n_pytest_rest.log
tail -6 gpurun_out/n_pytest_rest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/n_bench_c5.json 2> gpurun_out/n_bench_c5.err
echo "bench rc=$?"
THZ_CHAIN_SPECTRAL=off timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/n_bench_c5_plain.json 2> gpurun_out/n_bench_c5_plain.err
python - <<'PY'
import json
for c in ('c5','c5_plain'):
    try:
        d=json.loads(open(f'gpurun_out/n_bench_{c}.json').read().strip().splitlines()[-1])
        print(c,'ms_per_step %.3f'%d['ms_per_step'],'value %.3e'%d['value'], {k:round(v.get('ms'),2) for k,v in d['stage_breakdown'].items() if isinstance(v,dict) and 'ms' in v}, d['stage_breakdown'].get('stage_totals_ms'))
    except Exception as ex: print(c,'failed',ex)
PY
tail -3 gpurun_out/n_bench_c5.err
