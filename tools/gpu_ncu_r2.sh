#!/bin/bash
# Round-2 ncu evidence for the bench's workloads (each command first exits 0 WITHOUT ncu). Reports are summarised ON THE BOX
# (tools/ncu_summary.py) and deleted: gpurun copies back at most 64 MiB.
#   launch lists (gpu__time_duration.sum, --clock-control none) of the PICNet-ref batch-4 forward and the RefpSp batch-8 forward,
#   --set full of every implicit-GEMM launch + the attention kernels of one PICNet-ref forward (third forward of one_picnet.py 4),
#   and of the implicit-GEMM launches of one RefpSp forward.
mkdir -p gpurun_out
python tools/debug/one_picnet.py 4 fp32 > gpurun_out/r02_one_picnet.log 2>&1 || exit 1
python tools/debug/one_refpsp.py 8 bf16 > gpurun_out/r02_one_refpsp.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_picnet_b4.csv \
  python tools/debug/one_picnet.py 4 fp32 > gpurun_out/r02_ncu_l1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_refpsp_b8.csv \
  python tools/debug/one_refpsp.py 8 bf16 > gpurun_out/r02_ncu_l2.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:modconv_gemm_kernel --launch-skip 158 --launch-count 79 \
  -f -o /tmp/r02_prof_picnet_gemm python tools/debug/one_picnet.py 4 fp32 > gpurun_out/r02_ncu_f1.log 2>&1
python tools/ncu_summary.py /tmp/r02_prof_picnet_gemm.ncu-rep > gpurun_out/r02_ncu_picnet_gemm_summary.csv
timeout 300 ncu --set full --import-source on --clock-control none -k regex:attn_fwd2_kernel --launch-skip 5 --launch-count 1 \
  -f -o gpurun_out/r02_prof_picnet_attn python tools/debug/one_picnet.py 4 fp32 > gpurun_out/r02_ncu_f2.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_prof_picnet_attn.ncu-rep > gpurun_out/r02_ncu_picnet_attn_summary.csv
timeout 900 ncu --set full --clock-control none -k regex:modconv_gemm_kernel --launch-skip 208 --launch-count 104 \
  -f -o /tmp/r02_prof_refpsp_gemm python tools/debug/one_refpsp.py 8 bf16 > gpurun_out/r02_ncu_f3.log 2>&1
python tools/ncu_summary.py /tmp/r02_prof_refpsp_gemm.ncu-rep > gpurun_out/r02_ncu_refpsp_gemm_summary.csv
du -sh gpurun_out
