#!/bin/bash
# Round-1f GPU pass: loss tests, per-launch traces + one ncu metric pass of the general cuboid-attention kernel on
# non-axial patterns, final bench line, smoke.   gpurun -- bash tools/r01f_gpu.sh
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 120 python -m pytest tests/test_losses_gpu.py -q -p no:cacheprovider 2>&1 | tail -3
for p in video_swin_2x8,video_swin_2x8 divided_st,spatial_lg_4 full,axial_space_dilate_2; do
  timeout 100 python tools/trace_unet.py --batch 4 --graph --patterns $p --out $OUT/trace_r01f_unet_fwd_b4_${p//,/+}.txt | head -12
done
M='gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__shared_mem_per_block_dynamic,smsp__inst_executed.sum'
timeout 150 ncu --profile-from-start off --metrics $M --clock-control none --csv -k regex:cuboid_attention -c 6 \
    --log-file $OUT/ncu_r01f_cuboid_attention_b4.csv python tools/profile_unet.py --batch 4 --depth 1,1 \
    --patterns video_swin_2x8,divided_st > /dev/null 2>&1
timeout 150 ncu --profile-from-start off --metrics $M --clock-control none --csv -k regex:cuboid_attention -c 2 \
    --log-file $OUT/ncu_r01f_cuboid_attention_full_b4.csv python tools/profile_unet.py --batch 4 --depth 1,1 \
    --patterns full,spatial_lg_4 > /dev/null 2>&1
timeout 280 python bench.py > $OUT/bench_r01f.json 2> $OUT/bench_r01f.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r01f.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d.get('knowledge_alignment', {}).get('value'), d['roofline']['frac'], d['cpu_baseline']['value'], d['clocks'])
PY
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
