#!/bin/bash
# One GPU-box session for kernel experiments: parity suite, per-kernel A/B timings under the env switches given as
# arguments ("NAME=VALUE NAME=VALUE" per variant, "-" = defaults), sanitizer runs of a small case.  Output: gpurun_out/ab_*.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-ab}
mkdir -p rust-compression_b200/build/variants; cp rust-compression_b200/libbzb200.so rust-compression_b200/build/variants/lib_base.so
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gputests.log 2>&1; echo "tests rc $?"; tail -5 gpurun_out/${TAG}_gputests.log
: > gpurun_out/${TAG}_prof.log
for v in "$@"; do
  echo "== variant: $v" >> gpurun_out/${TAG}_prof.log
  if [ "$v" = "-" ]; then v=""; fi
  cp rust-compression_b200/build/variants/lib_base.so rust-compression_b200/libbzb200.so 2>/dev/null
  case "$v" in LIB=*) cp rust-compression_b200/build/variants/lib_${v#LIB=}.so rust-compression_b200/libbzb200.so; v="";; esac
  env $v timeout 300 python tools/gpu_enc_prof.py 1024 9 text >> gpurun_out/${TAG}_prof.log 2>&1
  env $v timeout 300 python tools/gpu_enc_prof.py 1024 1 mixed >> gpurun_out/${TAG}_prof.log 2>&1
done
cp rust-compression_b200/build/variants/lib_base.so rust-compression_b200/libbzb200.so
cat gpurun_out/${TAG}_prof.log
if [ -n "$LAUNCHES" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_1gib.csv \
      python tools/gpu_enc_once.py 1024 9 text > gpurun_out/${TAG}_launches.log 2>&1
fi
if [ -n "$SANITIZE" ]; then
  timeout 600 compute-sanitizer --tool memcheck python tools/gpu_enc_once.py 3 9 text > gpurun_out/${TAG}_memcheck.log 2>&1; tail -3 gpurun_out/${TAG}_memcheck.log
  timeout 600 compute-sanitizer --tool racecheck python tools/gpu_enc_once.py 3 9 text > gpurun_out/${TAG}_racecheck.log 2>&1; tail -3 gpurun_out/${TAG}_racecheck.log
fi
if [ -n "$NCU_K" ]; then
  ncu --set full --clock-control none --kernel-name-base demangled --kernel-id :::1 -k "regex:$NCU_K" -o /tmp/${TAG}_full \
      python tools/gpu_enc_once.py 256 9 text > gpurun_out/${TAG}_ncu_full.log 2>&1
  python profiles/summarize_ncu.py /tmp/${TAG}_full.ncu-rep gpurun_out/${TAG}_ncu_kernels.csv
  ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_raw_256mib.csv 2>/dev/null
fi
