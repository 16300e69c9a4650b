#!/bin/bash
# smoke + GPU tests + bench (both arms) on the GPU box; logs to gpurun_out/
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 100 --warmup 5 2> gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
python bench.py --impl reference --steps 10 --warmup 1 2>> gpurun_out/bench.err | tee gpurun_out/bench_ref.json
