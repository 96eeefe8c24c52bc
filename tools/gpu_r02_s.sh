#!/bin/bash
# round 2, call S (1 GPU): one barrier for the mirrors of both sub-spectra in k_chain_energy_fused
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chain_fused_gpu.py tests/test_trace_gpu.py -m gpu -q -x > gpurun_out/s_pytest.log 2>&1
tail -5 gpurun_out/s_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/s_bench_c5.json 2> gpurun_out/s_bench_c5.err
python - <<'PY'
import json
for c in ('c5',):
    try:
        d=json.loads(open(f'gpurun_out/s_bench_{c}.json').read().strip().splitlines()[-1])
        print(c,'ms_per_step %.3f'%d['ms_per_step'], {k:round(v.get('ms'),3) for k,v in d['stage_breakdown'].items() if isinstance(v,dict) and 'ms' in v})
    except Exception as ex: print(c,'failed',ex)
PY
