#!/bin/bash
# tcgen05 attention iteration: its tests + per-launch traces of the non-axial patterns (compare with trace_r02b_mma_*).
set -u
OUT=gpurun_out
TAG=${1:-r02c}
mkdir -p $OUT
timeout 300 python -m pytest tests/test_patterns_gpu.py -m gpu -x -q 2>&1 | tail -4
for p in video_swin_2x8,video_swin_2x8 divided_st,spatial_lg_4 full,axial_space_dilate_2 axial,full; do
  timeout 100 python tools/trace_unet.py --batch 4 --graph --patterns $p --out $OUT/trace_${TAG}_tc_${p//,/+}.txt | grep -E "forward|attn_cuboid"
done
