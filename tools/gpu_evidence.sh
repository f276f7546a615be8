#!/bin/bash
# GPU evidence pass for profiles/: per-site ncu metrics of one UNet forward (DRAM / L2 bytes, tensor-pipe %), the graph trace,
# the bench line, measured parity numbers, the ncu launch list of the bench command, smoke().
#   gpurun --timeout 1500 -- bash tools/gpu_evidence.sh r02m
set -u
TAG=${1:-r02m}
OUT=gpurun_out
mkdir -p $OUT
M='gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__shared_mem_per_block_dynamic'
python tools/dump_unet_labels.py 4 > $OUT/unet_labels_${TAG}_b4.txt
timeout 400 ncu --profile-from-start off --metrics $M --clock-control none --csv \
    --log-file $OUT/ncu_${TAG}_unet_fwd_b4.csv python tools/profile_unet.py --batch 4 > /dev/null 2>&1
timeout 100 python tools/trace_unet.py --batch 4 --graph --out $OUT/trace_${TAG}_unet_fwd_b4_graph.txt > /dev/null 2>&1
head -14 $OUT/trace_${TAG}_unet_fwd_b4_graph.txt
timeout 400 python -m pytest tests/test_tf32_gpu.py tests/test_sampler_round2_gpu.py tests/test_unet_gpu.py tests/test_sampler_gpu.py tests/test_global_vectors_gpu.py -m gpu -q -s 2>&1 | grep -E "rel_rms|passed|failed|ssim" > $OUT/parity_${TAG}.txt; tail -3 $OUT/parity_${TAG}.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_${TAG}_b200x1.json 2> $OUT/bench_${TAG}_b200x1.err; tail -2 $OUT/bench_${TAG}_b200x1.err
python -c "
import json; d=json.loads(open('$OUT/bench_${TAG}_b200x1.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d.get('tf32_arm',{}).get('value'), d.get('single_step_b1',{}).get('ms'), d.get('knowledge_alignment',{}).get('value'), d['cpu_baseline']['value'], d.get('gpu_launches'))
for k in d['roofline'].get('kernels', [])[:8]: print(k)
"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/launches_${TAG}_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ka --no-tf32 --no-extras > /dev/null 2>&1
python tools/summarize_launches.py $OUT/launches_${TAG}_bench.csv 2>/dev/null | head -20
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
