#!/bin/bash
# ncu evidence for profiles/: launch lists (device time per launch) + one `--set full` pass over selected kernels,
# exported as a small CSV of the metrics the roofline discussion uses (the .ncu-rep files are too big to ship back).
#   gpurun -- bash tools/ncu_capture.sh r01b
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
M='gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__shared_mem_per_block_dynamic,smsp__inst_executed.sum'
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $OUT/launches_${TAG}_unet_fwd_b4.csv python tools/profile_unet.py --batch 4 > /dev/null 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $OUT/launches_${TAG}_ka_b4.csv python tools/profile_ka.py --batch 4 > /dev/null 2>&1
# UNet: one level-0 resblock + one stack layer (launches 15..40 of the forward)
ncu --profile-from-start off --metrics $M --clock-control none --csv -s 14 -c 30 \
    --log-file $OUT/ncu_${TAG}_unet_block_b4.csv python tools/profile_unet.py --batch 4 > /dev/null 2>&1
# KA: the whole backward of level 1 + read-out head (first 60 launches after the forward's ~65)
ncu --profile-from-start off --metrics $M --clock-control none --csv -s 60 -c 60 \
    --log-file $OUT/ncu_${TAG}_ka_bwd_b4.csv python tools/profile_ka.py --batch 4 > /dev/null 2>&1
ls -la $OUT | tail -8
