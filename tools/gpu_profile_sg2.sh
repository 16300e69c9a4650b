#!/bin/bash
# ncu --set full of the two heaviest kernels of ONE StyleGAN2-1024 forward (bf16, batch 8): the last implicit-GEMM launch
# (32->32 @1024^2, launch 40 of 41) and the last blur+activation launch (1024^2, launch 7 of 8). Small reports only.
mkdir -p gpurun_out
python tools/debug/one_sg2.py || exit 1
ncu --set full --clock-control none --import-source on -k regex:modconv_gemm_kernel -s 40 -c 1 -f -o gpurun_out/prof_sg2_gemm \
    python tools/debug/one_sg2.py > gpurun_out/ncu_sg2_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:blur_act_nhwc -s 7 -c 1 -f -o gpurun_out/prof_sg2_blur \
    python tools/debug/one_sg2.py > gpurun_out/ncu_sg2_blur.log 2>&1
tail -2 gpurun_out/ncu_sg2_gemm.log gpurun_out/ncu_sg2_blur.log
ls -la gpurun_out/
