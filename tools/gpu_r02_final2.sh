#!/bin/bash
# round 2, final 1-GPU call of the final build: GPU suite, config-5 bench (default flags), reference arm, configs 1-4, ncu evidence
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/y_pytest_full.log 2>&1
echo "pytest rc=$?" >> gpurun_out/y_pytest_full.log
grep -E "passed|failed|rel err|RL band|config 3|FAILED|worst" gpurun_out/y_pytest_full.log | tail -12
python bench.py > gpurun_out/y_bench_c5.json 2> gpurun_out/y_bench_c5.err
echo "bench rc=$?" >> gpurun_out/y_bench_c5.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/y_bench_reference_arm.json 2> gpurun_out/y_bench_reference_arm.err
for c in c1 c2 c3 c4; do
  python bench.py --config $c --steps 5 --warmup 3 --no-cpu > gpurun_out/y_bench_$c.json 2> gpurun_out/y_bench_$c.err
  echo "$c rc=$?"
done
python - <<'PY'
import json
for c in ('c5','c1','c2','c3','c4'):
    try:
        d=json.loads(open(f'gpurun_out/y_bench_{c}.json').read().strip().splitlines()[-1])
        print(c,'ms_per_step %.3f'%d['ms_per_step'],'value %.3e'%d['value'],'e2e',(d.get('e2e') or {}).get('value'), {k:round(v.get('ms'),2) for k,v in d['stage_breakdown'].items() if isinstance(v,dict) and 'ms' in v})
    except Exception as ex: print(c,'failed',ex)
PY
# --- ncu evidence (the same commands have just exited 0 without ncu) ---
python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/y_plain_for_ncu.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_fir|k_trace|k_chain" -c 10 --csv \
    --log-file gpurun_out/y_ncu_cube_kernels_c5.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/y_ncu1.log 2>&1
python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --width 512 --height 512 > gpurun_out/y_plain512.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/y_ncu_launches_512.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --width 512 --height 512 > gpurun_out/y_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_chain_energy_fused|k_fir_edges_mma|k_fir_apply_circ|k_fir_edge_corr" -s 4 -c 4 \
    -o gpurun_out/y_prof_cube512 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --width 512 --height 512 > gpurun_out/y_ncu3.log 2>&1
python bench.py --config c4 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/y_plain_c4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_trace_fused" -s 2 -c 1 \
    -o gpurun_out/y_prof_trace_c4 python bench.py --config c4 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/y_ncu4.log 2>&1
ls -la gpurun_out/y_* | tail -24
