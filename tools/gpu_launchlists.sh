#!/bin/bash
# ncu launch lists of ONE forward (NVTX range "measured" of tools/debug/one_*.py), each after the same command exited 0 without ncu:
#   per-launch time + DRAM bytes (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none)
#   of the PICNet-ref batch-4 forward and the RefpSp batch-8 forward, and one --set full capture of the largest implicit-GEMM
#   launch (merged convT to 1024^2) and of the Auto_Attn kernel. Reports are summarised on the box. Usage: tools/gpu_launchlists.sh TAG
TAG=${1:-r02_final}
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
python tools/debug/one_picnet.py 4 fp32 > gpurun_out/${TAG}_one_picnet.log 2>&1 || exit 1
python tools/debug/one_refpsp.py 8 bf16 > gpurun_out/${TAG}_one_refpsp.log 2>&1 || exit 1
timeout 600 ncu --nvtx --nvtx-include "measured/" --metrics $M --clock-control none --csv --log-file gpurun_out/${TAG}_launches_picnet_b4.csv \
  python tools/debug/one_picnet.py 4 fp32 > gpurun_out/${TAG}_ncu_l1.log 2>&1
timeout 600 ncu --nvtx --nvtx-include "measured/" --metrics $M --clock-control none --csv --log-file gpurun_out/${TAG}_launches_refpsp_b8.csv \
  python tools/debug/one_refpsp.py 8 bf16 > gpurun_out/${TAG}_ncu_l2.log 2>&1
python tools/dram_summary.py gpurun_out/${TAG}_launches_picnet_b4.csv gpurun_out/${TAG}_traffic_picnet_b4.json > gpurun_out/${TAG}_launches_picnet_b4.txt
python tools/dram_summary.py gpurun_out/${TAG}_launches_refpsp_b8.csv gpurun_out/${TAG}_traffic_refpsp_b8.json 71 > gpurun_out/${TAG}_launches_refpsp_b8.txt
python tools/debug/one_picnet_conv.py 4 96 32 512 512 3 > gpurun_out/${TAG}_one_conv.log 2>&1 || exit 1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:modconv_gemm_kernel --launch-skip 2 --launch-count 1 -f -o /tmp/${TAG}_convT \
  python tools/debug/one_picnet_conv.py 4 96 32 512 512 3 > gpurun_out/${TAG}_ncu_f1.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_convT.ncu-rep > gpurun_out/${TAG}_ncu_merged_convT_summary.csv
ncu -i /tmp/${TAG}_convT.ncu-rep --page details > gpurun_out/${TAG}_ncu_merged_convT_details.txt 2>/dev/null
timeout 300 ncu --set full --clock-control none --nvtx --nvtx-include "measured/" -k regex:attn_fwd2_kernel --launch-skip 1 --launch-count 1 -f -o /tmp/${TAG}_attn \
  python tools/debug/one_picnet.py 4 fp32 > gpurun_out/${TAG}_ncu_f2.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_attn.ncu-rep > gpurun_out/${TAG}_ncu_picnet_attn_summary.csv
du -sh gpurun_out
