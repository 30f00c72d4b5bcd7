#!/bin/bash
# ncu --set full captures (with source counters) of single kernels, one launch each -> gpurun_out/ncu_*.ncu-rep
# usage: tools/ncu_batch.sh <tag>   (run on the GPU box)
TAG=${1:-r02}
NCU="ncu --set full --import-source on --clock-control none --launch-skip 1 --launch-count 1 -f"
mkdir -p gpurun_out
$NCU -k regex:old_blur_staged -o gpurun_out/ncu_${TAG}_blur_h_k7 python tools/blur_one.py h 0.0285 > /dev/null 2>&1
$NCU -k regex:old_blur_blocked -o gpurun_out/ncu_${TAG}_blur_h_k51 python tools/blur_one.py h 0.201 > /dev/null 2>&1
$NCU -k regex:old_blur -o gpurun_out/ncu_${TAG}_blur_v_k7 python tools/blur_one.py v 0.0285 > /dev/null 2>&1
$NCU -k regex:old_blur -o gpurun_out/ncu_${TAG}_blur_v_k51 python tools/blur_one.py v 0.201 > /dev/null 2>&1
$NCU -k regex:polar_blit_kernel -o gpurun_out/ncu_${TAG}_polar python tools/effect_one.py ball > /dev/null 2>&1
$NCU -k regex:landscape_kernel -o gpurun_out/ncu_${TAG}_landscape python tools/effect_one.py landscape > /dev/null 2>&1
$NCU -k regex:SpikeyDistant -o gpurun_out/ncu_${TAG}_spikey_distant python tools/effect_one.py spikey_distant > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
