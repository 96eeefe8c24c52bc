#!/bin/bash
# round 2, call V (1 GPU): A/B of the stash of k_chain_energy_fused in an L2-resident scratch (THZ_CHAIN_STASH=global)
mkdir -p gpurun_out
THZ_CHAIN_STASH=global timeout 600 python -m pytest tests/test_chain_fused_gpu.py -m gpu -q -x -k "spectral or two_passes" > gpurun_out/v_pytest.log 2>&1
tail -3 gpurun_out/v_pytest.log
THZ_CHAIN_STASH=global timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/v_bench_global.json 2> gpurun_out/v_bench_global.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/v_bench_shared.json 2> gpurun_out/v_bench_shared.err
python - <<'PY'
import json
for c in ('global','shared'):
    try:
        d=json.loads(open(f'gpurun_out/v_bench_{c}.json').read().strip().splitlines()[-1])
        print(c,'ms_per_step %.3f'%d['ms_per_step'], {k:round(v.get('ms'),3) for k,v in d['stage_breakdown'].items() if isinstance(v,dict) and 'ms' in v})
    except Exception as ex: print(c,'failed',ex)
PY
