#!/bin/bash
# round 2, call A (1 GPU): GPU tests, fp32 issue rates, short bench of config 5
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/a_pytest.log
echo "pytest rc=$?" >> gpurun_out/a_pytest.log
python - > gpurun_out/a_fp32.log 2>&1 <<'PY'
import importlib
m = importlib.import_module("thz-image-explorer_b200")
with m.Context(0) as c:
    names = ["FFMA", "FFMA2", "FADD", "FADD2", "FMUL", "FMUL2"]
    for mode, nm in enumerate(names):
        r = c.fp32_rate(mode)
        print(f"{nm:6s} {r/1e12:8.2f} T lane-ops/s  ({r/148/1.965e9:6.1f} lanes/clk/SM at 1965 MHz)")
PY
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/a_bench_c5.json 2> gpurun_out/a_bench_c5.err
echo "bench rc=$?" >> gpurun_out/a_bench_c5.err
tail -5 gpurun_out/a_pytest.log; cat gpurun_out/a_fp32.log; tail -c 1500 gpurun_out/a_bench_c5.json; tail -5 gpurun_out/a_bench_c5.err
