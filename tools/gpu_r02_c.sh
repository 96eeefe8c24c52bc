#!/bin/bash
# round 2, call C (2 GPUs): NVLink halo exchange tests + bench at N = 2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/c_gpus.txt
timeout 900 python -m pytest tests/test_slab_gpu.py -m gpu -q -s -k "two_gpus" > gpurun_out/c_pytest_2gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c_pytest_2gpu.log
tail -25 gpurun_out/c_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu > gpurun_out/c_bench_2gpu.json 2> gpurun_out/c_bench_2gpu.err
echo "bench rc=$?" >> gpurun_out/c_bench_2gpu.err
tail -c 2500 gpurun_out/c_bench_2gpu.json; tail -8 gpurun_out/c_bench_2gpu.err
