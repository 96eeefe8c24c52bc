#!/bin/bash
# round 2, call G (4 GPUs): per-launch time stamps of the slab Richardson-Lucy (THZ_SLAB_TRACE)
mkdir -p gpurun_out
THZ_SLAB_TRACE=gpurun_out/g_trace timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus 4 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/g_bench_4gpu.json 2> gpurun_out/g_bench_4gpu.err
echo "bench rc=$?" >> gpurun_out/g_bench_4gpu.err
python - <<'PY'
import json,csv
d=json.loads(open('gpurun_out/g_bench_4gpu.json').read().strip().splitlines()[-1])
print('ms_per_step',d['ms_per_step'], d['rank0_phases_ms'])
for r in range(4):
    rows=list(csv.DictReader(open(f'gpurun_out/g_trace.rank{r}.csv')))
    import statistics as st
    dur=[int(x['end_ns'])-int(x['start_ns']) for x in rows]
    gap=[int(rows[i+1]['start_ns'])-int(rows[i]['end_ns']) for i in range(len(rows)-1)]
    per=[int(rows[i+1]['start_ns'])-int(rows[i]['start_ns']) for i in range(len(rows)-1)]
    wait=[int(x['halo_wait_ns']) for x in rows]
    bend=[int(x['boundary_end_ns'])-int(x['start_ns']) for x in rows]
    def q(v,lo,hi): return st.median(v[lo:hi])
    for lo,hi in ((4,40),(50,120),(140,480),(520,840)):
        print(f'rank {r} launches {lo}-{hi}: period {q(per,lo,hi)/1e3:.1f} us, kernel {q(dur,lo,hi)/1e3:.1f} us, gap {q(gap,lo,hi)/1e3:.1f} us, halo wait {q(wait,lo,hi)/1e3:.1f} us, boundary done after {q(bend,lo,hi)/1e3:.1f} us')
PY
tail -3 gpurun_out/g_bench_4gpu.err
