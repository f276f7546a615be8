#!/bin/bash
# Round-2h GPU pass: measured parity numbers (pytest -s of the TF32 / round-2 sampler tests), the bench line, ncu metric pass
# of the tcgen05 attention kernel, smoke().
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests/test_tf32_gpu.py tests/test_sampler_round2_gpu.py tests/test_unet_gpu.py tests/test_sampler_gpu.py -m gpu -q -s 2>&1 | grep -E "rel_rms|passed|failed|ssim" > $OUT/parity_r02.txt; cat $OUT/parity_r02.txt
timeout 500 python bench.py --steps 20 --warmup 5 > $OUT/bench_r02_b200x1.json 2> $OUT/bench_r02_b200x1.err; tail -2 $OUT/bench_r02_b200x1.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r02_b200x1.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['tf32_arm']['value'], d['single_step_b1']['ms'], d['knowledge_alignment']['value'], d['cpu_baseline']['value'])
for k in d['roofline']['kernels'][:8]: print(k)
"
M='gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__shared_mem_per_block_dynamic,smsp__inst_executed.sum'
timeout 150 ncu --profile-from-start off --metrics $M --clock-control none --csv -k regex:cuboid_attention -c 6 \
    --log-file $OUT/ncu_r02_cuboid_attention_b4.csv python tools/profile_unet.py --batch 4 --depth 1,1 \
    --patterns divided_st,video_swin_2x8 > /dev/null 2>&1
timeout 150 ncu --profile-from-start off --metrics $M --clock-control none --csv -k regex:cuboid_attention -c 2 \
    --log-file $OUT/ncu_r02_cuboid_attention_full_b4.csv python tools/profile_unet.py --batch 4 --depth 1,1 \
    --patterns full,full > /dev/null 2>&1
python tools/pivot_ncu.py $OUT/ncu_r02_cuboid_attention_b4.csv; python tools/pivot_ncu.py $OUT/ncu_r02_cuboid_attention_full_b4.csv
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
