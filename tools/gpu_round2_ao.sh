#!/bin/bash
# work-proportional CTAs in the 32-channel row-streaming weight gradient: conv parity subset + the dominant kernel's ncu capture
mkdir -p gpurun_out
P=gpurun_out/r2l2
timeout 120 python -m pytest tests -m gpu -q -x -k "tensor_core_conv_matches_oracle or tall_conv or block_tail" > ${P}_sub.log 2>&1; echo "subset rc=$?" >> ${P}_sub.log
tail -2 ${P}_sub.log
KEY="cpc_conv_dgrad b64 128x63x156->128x34x156 k30x1 s1x1 [tall_conv_tcgen05_128ch]"
timeout 120 ncu --set full --clock-control none --import-source on -k regex:tall128_conv_kernel -c 2 -o ${P}_dominant python tools/profile_kernel.py "$KEY" 1 > ${P}_ncu_dom.log 2>&1
tail -2 ${P}_ncu_dom.log
