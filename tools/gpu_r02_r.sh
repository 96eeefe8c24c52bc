#!/bin/bash
# round 2, call R (1 GPU): packed adds by default + next-pair loads in flight across the tail of an item
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chain_fused_gpu.py tests/test_trace_gpu.py tests/test_deconv_gpu.py tests/test_chain_driver_gpu.py -m gpu -q -x --deselect tests/test_deconv_gpu.py::test_config3_full_chain_and_deconvolution_matches_oracle > gpurun_out/r_pytest.log 2>&1
tail -5 gpurun_out/r_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r_bench_c5.json 2> gpurun_out/r_bench_c5.err
for c in c4 c1; do timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r_bench_$c.json 2> gpurun_out/r_bench_$c.err; done
python - <<'PY'
import json
for c in ('c5','c4','c1'):
    try:
        d=json.loads(open(f'gpurun_out/r_bench_{c}.json').read().strip().splitlines()[-1])
        print(c,'ms_per_step %.3f'%d['ms_per_step'], {k:round(v.get('ms'),3) for k,v in d['stage_breakdown'].items() if isinstance(v,dict) and 'ms' in v})
    except Exception as ex: print(c,'failed',ex)
PY
