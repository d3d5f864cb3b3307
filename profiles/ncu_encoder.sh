#!/bin/bash
# ncu evidence for the encode kernels (run on the GPU box through gpurun; outputs under gpurun_out/, summaries only:
# the .ncu-rep files are tens of MB and are deleted after `ncu -i` has turned them into CSV on the box).
#  (1) launch list of one device-resident compression of the 1 GiB text corpus: device time of every launch
#      (cold-cache, serialised: compare shares with bench.py's kernels_ms_per_step, not absolutes)
#  (2) one `--set full` record per kernel (first launch of every kernel = the full-size one) on a 256 MiB corpus:
#      per-kernel table (profiles/summarize_ncu.py) + the raw page of every metric
#  (3) the same for the sort kernels on the level-1 mixed corpus (byte-key sort: k2_os_scatter<8,5,8>, k2_os_hist_txt)
TAG=${1:-r2}
export BZB200_TRACE_ROUNDS=1   # per-round work of the sort in the logs: the algorithmic bytes of the profiled launches
cd "$(dirname "$0")/.."
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_1gib.csv \
    python tools/gpu_enc_once.py 1024 9 text > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --kernel-name-base demangled --kernel-id :::1 -o /tmp/${TAG}_full \
    python tools/gpu_enc_once.py 256 9 text > gpurun_out/${TAG}_ncu_full.log 2>&1
python profiles/summarize_ncu.py /tmp/${TAG}_full.ncu-rep gpurun_out/${TAG}_ncu_kernels.csv
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_raw_256mib.csv 2>/dev/null
# DRAM bytes next to the algorithmic bytes of the same launches: the table bench.py reads for roofline.traffic
python profiles/make_traffic_table.py gpurun_out/${TAG}_ncu_raw_256mib.csv gpurun_out/${TAG}_ncu_full.log gpurun_out/${TAG}_ncu_traffic.csv
ncu --set full --clock-control none --kernel-name-base demangled --kernel-id :::1 -k 'regex:k2_os|k2_pair|k2_digit' -o /tmp/${TAG}_mixed \
    python tools/gpu_enc_once.py 128 1 mixed > gpurun_out/${TAG}_ncu_mixed.log 2>&1
python profiles/summarize_ncu.py /tmp/${TAG}_mixed.ncu-rep gpurun_out/${TAG}_ncu_kernels_mixed.csv
rm -f /tmp/${TAG}_full.ncu-rep /tmp/${TAG}_mixed.ncu-rep
du -sh gpurun_out | tail -1
