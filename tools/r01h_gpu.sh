#!/bin/bash
# Round-1h GPU pass: the whole GPU suite on the final state, then traces of the two patterns whose schedule changed
# (full attention -> single stage + global bias gather; volume-4 dilated cuboids -> single stage).
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > $OUT/pytest_r01h.log 2>&1; tail -6 $OUT/pytest_r01h.log
timeout 60 python tools/trace_unet.py --batch 4 --graph --patterns full,axial_space_dilate_2 --out $OUT/trace_r01h_unet_fwd_b4_full+axial_space_dilate_2.txt | grep -E "UNet forward|attn_cuboid"
timeout 60 python tools/trace_unet.py --batch 4 --graph --patterns video_swin_2x8,spatial_lg_4 --out $OUT/trace_r01h_unet_fwd_b4_video_swin_2x8+spatial_lg_4.txt | grep -E "UNet forward|attn_cuboid"
