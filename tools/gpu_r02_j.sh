#!/bin/bash
# round 2, call J (8 GPUs): slab tests, then bench at 8 / 4 / 2 GPUs with the two-table segmentation
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_slab_gpu.py -m gpu -q -x > gpurun_out/j_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j_pytest.log
tail -4 gpurun_out/j_pytest.log
for n in 8 4 2; do
THZ_SLAB_TRACE=gpurun_out/j_trace_${n} timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2956$n \
  bench.py --gpus $n --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/j_bench_${n}gpu.json 2> gpurun_out/j_bench_${n}gpu.err
echo "bench rc=$?" >> gpurun_out/j_bench_${n}gpu.err
python - $n <<'PY'
import json,csv,sys
import statistics as st
n=int(sys.argv[1])
d=json.loads(open(f'gpurun_out/j_bench_{n}gpu.json').read().strip().splitlines()[-1])
print(n,'GPUs ms_per_step',d['ms_per_step'], d['rank0_phases_ms'])
for r in (0, n//2):
    rows=list(csv.DictReader(open(f'gpurun_out/j_trace_{n}.rank{r}.csv')))
    dur=[int(x['end_ns'])-int(x['start_ns']) for x in rows]
    per=[int(rows[i+1]['start_ns'])-int(rows[i]['start_ns']) for i in range(len(rows)-1)]
    wait=[int(x['halo_wait_ns']) for x in rows]
    bend=[int(x['boundary_end_ns'])-int(x['start_ns']) for x in rows]
    def q(v,lo,hi): return st.median(v[lo:hi])
    out=[]
    for lo,hi in ((4,26),(30,90),(100,250),(260,500),(520,840)):
        out.append(f'[{lo}-{hi}] period {q(per,lo,hi)/1e3:.1f} kernel {q(dur,lo,hi)/1e3:.1f} wait {q(wait,lo,hi)/1e3:.1f} bdone {q(bend,lo,hi)/1e3:.1f}')
    print('rank',r,' | '.join(out))
PY
tail -2 gpurun_out/j_bench_${n}gpu.err
done
