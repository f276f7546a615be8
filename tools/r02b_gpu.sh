#!/bin/bash
# Round-2b GPU pass: full GPU test suite, per-launch traces of the non-axial patterns with the tcgen05 attention tile vs the
# mma.sync kernel (PD_CUBOID_NO_TC=1), one ncu metric pass of the tcgen05 kernel.   gpurun -- bash tools/r02b_gpu.sh
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $OUT/pytest_r02b.log; tail -4 $OUT/pytest_r02b.log
for p in video_swin_2x8,video_swin_2x8 divided_st,spatial_lg_4 full,axial_space_dilate_2 axial,full; do
  timeout 100 python tools/trace_unet.py --batch 4 --graph --patterns $p --out $OUT/trace_r02b_tc_${p//,/+}.txt | grep -E "forward|attn_cuboid"
  PD_CUBOID_NO_TC=1 timeout 100 python tools/trace_unet.py --batch 4 --graph --patterns $p --out $OUT/trace_r02b_mma_${p//,/+}.txt | grep -E "forward|attn_cuboid"
done
M='gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__shared_mem_per_block_dynamic,smsp__inst_executed.sum'
timeout 150 ncu --profile-from-start off --metrics $M --clock-control none --csv -k regex:cuboid_attention -c 6 \
    --log-file $OUT/ncu_r02b_cuboid_attention_tc_b4.csv python tools/profile_unet.py --batch 4 --depth 1,1 \
    --patterns video_swin_2x8,divided_st > /dev/null 2>&1
timeout 150 ncu --profile-from-start off --metrics $M --clock-control none --csv -k regex:cuboid_attention -c 2 \
    --log-file $OUT/ncu_r02b_cuboid_attention_tc_full_b4.csv python tools/profile_unet.py --batch 4 --depth 1,1 \
    --patterns full,full > /dev/null 2>&1
python tools/pivot_ncu.py $OUT/ncu_r02b_cuboid_attention_tc_b4.csv; python tools/pivot_ncu.py $OUT/ncu_r02b_cuboid_attention_tc_full_b4.csv
