#!/bin/bash
# per-kernel times of one bench step on a 512x512x4096 cube (1/16 of config 5), via the ncu launch list
python bench.py --no-cpu --no-e2e --steps 2 --warmup 1 --width 512 --height 512 > gpurun_out/b512.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_fir|k_trace" -c 10 --csv --log-file gpurun_out/launches_512.csv python bench.py --no-cpu --no-e2e --steps 2 --warmup 1 --width 512 --height 512 > gpurun_out/ncu512.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_512.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:11]: print(r[ki][:50], float(r[vi])/1e6, 'ms')
PY
