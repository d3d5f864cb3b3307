#!/bin/bash
# ncu evidence for the encode kernels (run on the GPU box through gpurun; outputs under gpurun_out/).
#  (1) launch list of one device-resident compression of the 1 GiB text corpus: device time of every launch
#      (cold-cache, serialised: compare shares with bench.py's kernels_ms_per_step, not absolutes)
#  (2) one `--set full` record per kernel (first launch of every kernel = the full-size one) on a 256 MiB corpus
#  (3) the same for the level-1 mixed corpus (byte-key sort: k2_os_scatter<8,5,8>, k2_os_hist_txt)
TAG=${1:-r2}
cd "$(dirname "$0")/.."
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_1gib.csv \
    python tools/gpu_enc_once.py 1024 9 text > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-id :::1 -o gpurun_out/${TAG}_ncu_full_256mib \
    python tools/gpu_enc_once.py 256 9 text > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu --set full --clock-control none --kernel-id :::1 -k 'regex:k2_os|k2_pair|k2_digit' -o gpurun_out/${TAG}_ncu_mixed_128mib \
    python tools/gpu_enc_once.py 128 1 mixed > gpurun_out/${TAG}_ncu_mixed.log 2>&1
ls -la gpurun_out/ | tail -8
