#!/bin/bash
# Round-2f GPU pass: ncu evidence for the UNet (per-site DRAM bytes) and the VAE (HBM GB/s of the GN+SiLU kernels, tensor
# pipe of the convs / attention), compute-sanitizer memcheck / racecheck over the op-level tests.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 200 python -m pytest tests/test_patterns_gpu.py -m gpu -x -q 2>&1 | tail -2
M='gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__shared_mem_per_block_dynamic'
python tools/dump_unet_labels.py 4 > $OUT/unet_labels_r02_b4.txt
timeout 400 ncu --profile-from-start off --metrics $M --clock-control none --csv \
    --log-file $OUT/ncu_r02_unet_fwd_b4.csv python tools/profile_unet.py --batch 4 > /dev/null 2>&1
timeout 400 ncu --profile-from-start off --metrics $M --clock-control none --csv \
    --log-file $OUT/ncu_r02_vae_b4.csv python tools/profile_vae.py --profile > /dev/null 2>&1
python tools/profile_vae.py | tee $OUT/vae_time_r02.txt
ls -la $OUT/ncu_r02_*.csv
# sanitizers on the op-level tests (small shapes); racecheck sees shared-memory hazards only
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_ops_gpu.py tests/test_tf32_gpu.py -m gpu -x -q \
    -k "not full and not ddim50" -p no:cacheprovider > $OUT/sanitizer_r02_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $OUT/sanitizer_r02_memcheck.log
tail -5 $OUT/sanitizer_r02_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_ops_gpu.py -m gpu -x -q \
    -k "layer_norm or group_norm or axial or patch_merge or sampler or upsample" -p no:cacheprovider > $OUT/sanitizer_r02_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $OUT/sanitizer_r02_racecheck.log
tail -5 $OUT/sanitizer_r02_racecheck.log
