#!/bin/bash
# ncu --set full captures of the PICNet-ref batch-8 forward (third forward of tools/debug/one_picnet.py): the last 12 implicit-GEMM
# launches (decoder blocks 2-4), the Output kernel, the last 6 InstanceNorm statistics / normalise+activate launches.
mkdir -p gpurun_out
python tools/debug/one_picnet.py 8 fp32 > gpurun_out/one_picnet.log 2>&1 || exit 1
timeout 500 ncu --set full --import-source on --clock-control none -k regex:modconv_gemm_kernel --launch-skip 225 --launch-count 12 \
  -f -o gpurun_out/prof_picnet_gemm python tools/debug/one_picnet.py 8 fp32 > gpurun_out/ncu_full_gemm.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:out_conv_tanh --launch-skip 2 --launch-count 1 \
  -f -o gpurun_out/prof_picnet_outconv python tools/debug/one_picnet.py 8 fp32 > gpurun_out/ncu_full_outconv.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"norm_act_kernel|instnorm_stats_kernel" --launch-skip 105 \
  --launch-count 6 -f -o gpurun_out/prof_picnet_stream python tools/debug/one_picnet.py 8 fp32 > gpurun_out/ncu_full_stream.log 2>&1
ls -la gpurun_out/*.ncu-rep
