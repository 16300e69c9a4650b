#!/bin/bash
# First GPU contact: sm_100a primitive probe (one process per test) + parity tests of the HBM-bound ops.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import os; print('cpu_count', os.cpu_count())" >> gpurun_out/gpu.txt
for t in 0 1 2 3 4 5 6 7 8; do
  timeout 60 ./build/umma_probe $t >> gpurun_out/probe.log 2>&1
  echo "test $t exit $?" >> gpurun_out/probe.log
done
cat gpurun_out/probe.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
