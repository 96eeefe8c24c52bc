#!/bin/bash
# round 2, call Q (1 GPU): A/B of packed FP32 add / subtract in the butterflies (THZ_PACKED_ADD build in _ab/)
mkdir -p gpurun_out
PK=thz-image-explorer_b200
cp $PK/libthzgpu.so /tmp/libthzgpu_default.so
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/q_bench_default.json 2> gpurun_out/q_bench_default.err
cp $PK/_ab/libthzgpu_packed.so $PK/libthzgpu.so
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/q_bench_packed.json 2> gpurun_out/q_bench_packed.err
timeout 900 python -m pytest tests/test_chain_fused_gpu.py tests/test_trace_gpu.py tests/test_deconv_gpu.py -m gpu -q -x --deselect tests/test_deconv_gpu.py::test_config3_full_chain_and_deconvolution_matches_oracle > gpurun_out/q_pytest_packed.log 2>&1
tail -5 gpurun_out/q_pytest_packed.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_chain_energy_fused" -s 1 -c 1 \
    -o gpurun_out/q_prof_chain_fused_packed python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --width 512 --height 512 > gpurun_out/q_ncu.log 2>&1
cp /tmp/libthzgpu_default.so $PK/libthzgpu.so
python - <<'PY'
import json
for c in ('default','packed'):
    try:
        d=json.loads(open(f'gpurun_out/q_bench_{c}.json').read().strip().splitlines()[-1])
        print(c,'ms_per_step %.3f'%d['ms_per_step'], {k:round(v.get('ms'),2) for k,v in d['stage_breakdown'].items() if isinstance(v,dict) and 'ms' in v})
    except Exception as ex: print(c,'failed',ex)
PY
