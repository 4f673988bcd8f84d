#!/bin/bash
mkdir -p gpurun_out
K1="cpc_conv_dgrad b64 128x34x156->256x16x77 k3x3 s2x2 [conv_implicit_gemm_tcgen05]"
K2="cpc_conv_dgrad b64 32x127x314->128x63x156 k3x3 s2x2 [conv_implicit_gemm_tcgen05]"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"umma_conv_kernel|umma_wgrad_kernel|pack" -c 40 -o gpurun_out/r2p_conv1 python tools/profile_kernel.py "$K1" 1 > gpurun_out/r2p_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"umma_conv_kernel|umma_wgrad_kernel|pack" -c 40 -o gpurun_out/r2p_conv2 python tools/profile_kernel.py "$K2" 1 > gpurun_out/r2p_ncu2.log 2>&1
tail -4 gpurun_out/r2p_ncu1.log
