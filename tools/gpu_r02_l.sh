#!/bin/bash
# round 2, call L (1 GPU): fused trace + band-energy kernel -- tests, config-5 bench, ncu of the new kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chain_fused_gpu.py -m gpu -q -x > gpurun_out/l_pytest_fused.log 2>&1
echo "pytest rc=$?" >> gpurun_out/l_pytest_fused.log
tail -15 gpurun_out/l_pytest_fused.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_chain_fused_gpu.py > gpurun_out/l_pytest_rest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/l_pytest_rest.log
tail -6 gpurun_out/l_pytest_rest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/l_bench_c5.json 2> gpurun_out/l_bench_c5.err
echo "bench rc=$?"
THZ_CHAIN_EVEN=transform timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/l_bench_c5_even_transform.json 2> gpurun_out/l_bench_c5_even_transform.err
python - <<'PY'
import json
for c in ('c5','c5_even_transform'):
    try:
        d=json.loads(open(f'gpurun_out/l_bench_{c}.json').read().strip().splitlines()[-1])
        print(c,'ms_per_step %.3f'%d['ms_per_step'],'value %.3e'%d['value'], {k:round(v.get('ms'),2) for k,v in d['stage_breakdown'].items() if isinstance(v,dict) and 'ms' in v}, d['stage_breakdown'].get('stage_totals_ms'))
    except Exception as ex: print(c,'failed',ex)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_chain_energy_fused" -s 2 -c 1 \
    -o gpurun_out/l_prof_chain_fused python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --width 512 --height 512 > gpurun_out/l_ncu.log 2>&1
ls -la gpurun_out/l_* | tail
