#!/bin/bash
# round 2, call B (1 GPU): the whole GPU suite
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s --durations=8 > gpurun_out/b_pytest_full.log 2>&1
echo "pytest rc=$?" >> gpurun_out/b_pytest_full.log
grep -E "passed|failed|rel err|RL band|config 3|FAILED|Error|error" gpurun_out/b_pytest_full.log | tail -40
