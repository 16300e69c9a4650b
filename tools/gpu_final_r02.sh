#!/bin/bash
# Round-2 closing run on one B200: smoke, the whole GPU test suite, the default bench (all four records), the streaming-kernel table and the
# torch.profiler breakdown of the GAN train step. Logs / JSON to gpurun_out/ (copied into profiles/ afterwards). Usage: tools/gpu_final_r02.sh [TAG]
TAG=${1:-r02_final2}
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -4 gpurun_out/${TAG}_pytest_gpu.log
python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; tail -2 gpurun_out/${TAG}_bench_n1.err
python - <<EOF
import json
d = json.loads([l for l in open("gpurun_out/${TAG}_bench_n1.json") if l.startswith("{")][-1])
for r in [d] + d.get("records", []):
    rf = r.get("roofline") or {}
    print(r["config"]["name"], round(r["ms_per_step"], 3), "ms", round(r["value"], 1), "img/s  e2e", round(r["e2e"]["value"], 1),
          "| roofline", rf.get("kernel"), rf.get("bound"), round(rf.get("frac", 0), 3), "sol", round(rf.get("sol_frac", 0), 3), "traffic", rf.get("traffic"))
EOF
python tools/perf/perf_streaming.py > gpurun_out/${TAG}_perf_streaming.txt 2>&1; grep upfirdn gpurun_out/${TAG}_perf_streaming.txt
python tools/debug/prof_train_picnet.py > gpurun_out/${TAG}_torchprof_train_picnet.txt 2>&1; tail -3 gpurun_out/${TAG}_torchprof_train_picnet.txt
