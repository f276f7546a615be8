#!/bin/bash
# Round-2j GPU pass: cluster FFN with the L2 reduce-scatter - op test, phase stamps, A/B bench on one box.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "ffn_cluster" 2>&1 | tail -15
timeout 120 python tools/ffn_cluster_phases.py 2>&1 | tee $OUT/ffn_cluster_phases_r02j.txt
timeout 300 python -m pytest tests/test_unet_gpu.py tests/test_sampler_gpu.py -m gpu -q -x 2>&1 | tail -5
source tools/ab.sh
run cluster A=1
run nocluster PD_NO_L1_FFN_FUSION=1
