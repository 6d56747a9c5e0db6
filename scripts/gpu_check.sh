#!/bin/bash
# Run on the GPU box: the whole GPU test-suite with a hard wall-clock bound.
mkdir -p gpurun_out
date +%s > gpurun_out/t0
timeout -s KILL ${1:-900} python -m pytest tests -q -m gpu -x -s --durations=12 ${3:-} 2>&1 | tail -${2:-60} | tee gpurun_out/gpu_tests.log
echo "elapsed $(( $(date +%s) - $(cat gpurun_out/t0) )) s"
