#!/bin/bash
# ncu evidence for the decoder kernels (run on the GPU box through gpurun; outputs under gpurun_out/).
#  (1) launch list of one decode of the 256 MiB stream: every decoder kernel, device time of each launch
#  (2) `--set full` records of the short kernels (64 MiB stream)
#  (3) d2_huff (the serial Huffman chain; d2_decode when BZB200_DEC_SPLIT=0): its duration is one block's serial chain whatever the block count, and ncu replays it ~40 times per
#      full set, so it gets the sections that explain a latency-bound kernel on an 8 MiB stream instead
cd "$(dirname "$0")/.."
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(d[0-9]_|k5_crc)' -c 14 --csv \
    --log-file gpurun_out/r1c_dec_launches.csv python tools/gpu_dec_bench.py 256 9 text > gpurun_out/r1c_dec_launches.log 2>&1
ncu --section SpeedOfLight --section WarpStateStats --section SchedulerStats --section Occupancy --section LaunchStats \
    --section InstructionStats --section MemoryWorkloadAnalysis --clock-control none -k 'regex:^d2_huff' -c 1 \
    -o gpurun_out/r1c_dec_ncu_d2 python tools/gpu_dec_bench.py 8 9 text > gpurun_out/r1c_dec_ncu_d2.log 2>&1
ncu --set full --clock-control none -k 'regex:^(d[13-5]_|d2_mtf)' -c 11 -o gpurun_out/r1c_dec_ncu_small \
    python tools/gpu_dec_bench.py 64 9 text > gpurun_out/r1c_dec_ncu_small.log 2>&1
ls -la gpurun_out/
