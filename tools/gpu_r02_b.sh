#!/usr/bin/env bash
# Round 2, GPU session B: full GPU suite (no -x), launch list + full ncu capture of the two render kernels (new addressing).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02b_tests.log 2>&1
echo "tests rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r02b_tests.log | tail -30
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'nff_|raygen|lidar_decode' -c 40 --csv --log-file gpurun_out/r02b_launches.csv \
  env IMAGE_WIDTH=640 python tools/perf_probe.py 0 3 > gpurun_out/r02b_ncu.log 2>&1
grep -E "nff_|raygen" gpurun_out/r02b_launches.csv | awk -F'","' '{print $5, $NF}' | tail -6
ncu --set full --clock-control none --import-source on -k regex:nff_s -s 4 -c 2 -o gpurun_out/r02b_prof_split \
  env IMAGE_WIDTH=640 python tools/perf_probe.py 0 1 >> gpurun_out/r02b_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
