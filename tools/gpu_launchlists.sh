#!/bin/bash
# ncu launch lists (gpu__time_duration.sum, --clock-control none) of one PICNet-ref batch-4 forward (TF32 and strict-fp32 split
# operands) and one RefpSp batch-8 forward, each after the same command exited 0 without ncu. Usage: tools/gpu_launchlists.sh TAG
TAG=${1:-r02_final}
mkdir -p gpurun_out
python tools/debug/one_picnet.py 4 fp32 > gpurun_out/${TAG}_one_picnet.log 2>&1 || exit 1
python tools/debug/one_refpsp.py 8 bf16 > gpurun_out/${TAG}_one_refpsp.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches_picnet_b4.csv \
  python tools/debug/one_picnet.py 4 fp32 > gpurun_out/${TAG}_ncu_l1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches_refpsp_b8.csv \
  python tools/debug/one_refpsp.py 8 bf16 > gpurun_out/${TAG}_ncu_l2.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches_picnet_b4.csv > gpurun_out/${TAG}_launches_picnet_b4.txt
python tools/launch_summary.py gpurun_out/${TAG}_launches_refpsp_b8.csv > gpurun_out/${TAG}_launches_refpsp_b8.txt
du -sh gpurun_out
