#!/bin/bash
# ncu --set full of the weight-gradient GEMM and the data-gradient implicit GEMM of one training convolution (64 -> 32 @512^2, batch 4),
# after the same command exited 0 without ncu. Usage: tools/gpu_ncu_train.sh [TAG]
TAG=${1:-r02}
mkdir -p gpurun_out
python tools/debug/one_conv_train.py > gpurun_out/${TAG}_one_conv_train.log 2>&1 || exit 1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:wgrad_gemm_kernel --launch-skip 2 --launch-count 1 -f -o /tmp/${TAG}_wgrad \
  python tools/debug/one_conv_train.py > gpurun_out/${TAG}_ncu_wgrad.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_wgrad.ncu-rep > gpurun_out/${TAG}_ncu_wgrad_summary.csv
ncu -i /tmp/${TAG}_wgrad.ncu-rep --page details > gpurun_out/${TAG}_ncu_wgrad_details.txt 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:modconv_gemm_kernel --launch-skip 4 --launch-count 2 -f -o /tmp/${TAG}_dgrad \
  python tools/debug/one_conv_train.py >> gpurun_out/${TAG}_ncu_wgrad.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_dgrad.ncu-rep > gpurun_out/${TAG}_ncu_conv_train_gemm_summary.csv
cat gpurun_out/${TAG}_ncu_wgrad_summary.csv gpurun_out/${TAG}_ncu_conv_train_gemm_summary.csv
grep -E "Duration|DRAM Throughput|Issue Slots Busy|Executed Ipc Active|Achieved Occupancy|L2 Hit|Mem Busy|No Eligible|Warp Cycles Per Issued|Registers Per" gpurun_out/${TAG}_ncu_wgrad_details.txt
