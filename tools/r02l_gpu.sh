#!/bin/bash
# Round-2l GPU pass: fused-FFN kernel changes - op tests, phase stamps, UNet / sampler parity, A/B against a saved build.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "ffn" 2>&1 | tail -5
timeout 120 python tools/ffn_phases.py 2>&1 | tail -7
timeout 120 python tools/ffn_cluster_phases.py 2>&1 | tail -3
timeout 400 python -m pytest tests/test_unet_gpu.py tests/test_sampler_gpu.py -m gpu -q -x 2>&1 | tail -3
source tools/ab.sh
run new A=1
run base PD_LIB_PATH=$PWD/prediff_b200/libprediff_b200_base.so
