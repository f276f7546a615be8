#!/bin/bash
# Round-1g GPU pass: pattern parity after the cuboid-attention rewrite (double-buffered chunks, bias column in shared
# memory) + the same three per-launch traces as r01f.   gpurun -- bash tools/r01g_gpu.sh
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 150 python -m pytest tests/test_patterns_gpu.py tests/test_losses_gpu.py -q --tb=short -p no:cacheprovider 2>&1 | tail -15
for p in video_swin_2x8,video_swin_2x8 divided_st,spatial_lg_4 full,axial_space_dilate_2; do
  timeout 100 python tools/trace_unet.py --batch 4 --graph --patterns $p --out $OUT/trace_r01g_unet_fwd_b4_${p//,/+}.txt | grep -E "UNet forward|attn_cuboid"
done
