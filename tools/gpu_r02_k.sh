#!/bin/bash
# round 2, call K (2 GPUs): fine time stamps inside one boundary CTA and one interior CTA of the slab RL kernels
mkdir -p gpurun_out
THZ_SLAB_TRACE=gpurun_out/k_trace timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 \
  bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu --no-e2e --width 512 > gpurun_out/k_bench_2gpu.json 2> gpurun_out/k_bench_2gpu.err
echo "bench rc=$?" >> gpurun_out/k_bench_2gpu.err
python - <<'PY'
import json,csv
import statistics as st
d=json.loads(open('gpurun_out/k_bench_2gpu.json').read().strip().splitlines()[-1])
print('ms_per_step',d['ms_per_step'], d['rank0_phases_ms'])
for r in (0,1):
    rows=list(csv.DictReader(open(f'gpurun_out/k_trace.rank{r}.csv')))
    def med(key,lo,hi): return st.median([int(x[key]) for x in rows[lo:hi]])/1e3
    for lo,hi in ((30,120),(520,840)):
        print(f'rank {r} [{lo}-{hi}] kernel {med("end_ns",lo,hi)-med("start_ns",lo,hi):.1f} us | top CTA: t0 {med("top_t0",lo,hi):.1f} wait_done {med("top_t1",lo,hi):.1f} chunk0 {med("top_t2",lo,hi):.1f} loop_done {med("top_t3",lo,hi):.1f} signalled {med("top_t4",lo,hi):.1f} | interior CTA: t0 {med("int_t0",lo,hi):.1f} t1 {med("int_t1",lo,hi):.1f} chunk0 {med("int_t2",lo,hi):.1f} loop_done {med("int_t3",lo,hi):.1f} end {med("int_t4",lo,hi):.1f}')
PY
tail -2 gpurun_out/k_bench_2gpu.err
