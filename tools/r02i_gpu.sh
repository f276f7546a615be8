#!/bin/bash
# Round-2i GPU pass: the width-512 cluster FFN kernel - op test, graph trace with / without it, A/B bench on one box.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "ffn_cluster" 2>&1 | tail -15
timeout 200 python tools/trace_unet.py --batch 4 --graph --out $OUT/trace_r02i_cluster.txt > /dev/null 2>$OUT/trace_r02i.err; head -30 $OUT/trace_r02i_cluster.txt
PD_NO_L1_FFN_FUSION=1 timeout 200 python tools/trace_unet.py --batch 4 --graph --out $OUT/trace_r02i_nocluster.txt > /dev/null 2>&1; head -30 $OUT/trace_r02i_nocluster.txt
timeout 300 python -m pytest tests/test_unet_gpu.py tests/test_sampler_gpu.py -m gpu -q -x 2>&1 | tail -5
source tools/ab.sh
run cluster A=1
run nocluster PD_NO_L1_FFN_FUSION=1
run cluster2 A=1
