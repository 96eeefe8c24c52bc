#!/bin/bash
# round 2, call M (1 GPU): fused-kernel tests, ncu --set full of k_chain_energy_fused on a 512 x 512 x 4096 cube
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chain_fused_gpu.py -m gpu -q > gpurun_out/m_pytest_fused.log 2>&1
echo "pytest rc=$?" >> gpurun_out/m_pytest_fused.log
tail -25 gpurun_out/m_pytest_fused.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_chain_energy_fused" -s 1 -c 1 \
    -o gpurun_out/m_prof_chain_fused python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --width 512 --height 512 > gpurun_out/m_ncu.log 2>&1
ls -la gpurun_out/m_* | tail
