#!/bin/bash
# Round-2k GPU pass: half2 GELU in the fused FFN epilogues - op tests, phase stamps, parity numbers, A/B against the previous build.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "ffn" 2>&1 | tail -5
timeout 120 python tools/ffn_phases.py 2>&1 | tail -7
timeout 120 python tools/ffn_cluster_phases.py 2>&1 | tail -20
timeout 400 python -m pytest tests/test_unet_gpu.py tests/test_sampler_gpu.py tests/test_sampler_round2_gpu.py -m gpu -q -s 2>&1 | grep -E "rel_rms|passed|failed" | tee $OUT/parity_r02k.txt
source tools/ab.sh
run new A=1
run base PD_LIB_PATH=$PWD/prediff_b200/libprediff_b200_base.so
