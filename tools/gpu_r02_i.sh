#!/bin/bash
# round 2, call I (8 GPUs): slab Richardson-Lucy after boundary-first ordering / balanced segments, with launch traces
mkdir -p gpurun_out
for n in 8 4; do
THZ_SLAB_TRACE=gpurun_out/i_trace_${n} timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n \
  bench.py --gpus $n --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/i_bench_${n}gpu.json 2> gpurun_out/i_bench_${n}gpu.err
echo "bench rc=$?" >> gpurun_out/i_bench_${n}gpu.err
python - $n <<'PY'
import json,csv,sys
import statistics as st
n=int(sys.argv[1])
d=json.loads(open(f'gpurun_out/i_bench_{n}gpu.json').read().strip().splitlines()[-1])
print(n,'GPUs ms_per_step',d['ms_per_step'], d['rank0_phases_ms'])
for r in range(n):
    rows=list(csv.DictReader(open(f'gpurun_out/i_trace_{n}.rank{r}.csv')))
    dur=[int(x['end_ns'])-int(x['start_ns']) for x in rows]
    per=[int(rows[i+1]['start_ns'])-int(rows[i]['start_ns']) for i in range(len(rows)-1)]
    wait=[int(x['halo_wait_ns']) for x in rows]
    bend=[int(x['boundary_end_ns'])-int(x['start_ns']) for x in rows]
    def q(v,lo,hi): return st.median(v[lo:hi])
    out=[]
    for lo,hi in ((4,40),(140,480),(520,840)):
        out.append(f'[{lo}-{hi}] period {q(per,lo,hi)/1e3:.1f} kernel {q(dur,lo,hi)/1e3:.1f} wait {q(wait,lo,hi)/1e3:.1f} bdone {q(bend,lo,hi)/1e3:.1f}')
    print('rank',r,' | '.join(out))
PY
tail -2 gpurun_out/i_bench_${n}gpu.err
done
