#!/usr/bin/env bash
# Round 2, GPU session S (final evidence for the committed sources): full GPU suite, training step + its launch list, ncu --set
# full of the two render kernels (profiles/traffic.json), launch list of bench.py's own command, N = 1 bench line, reference arm.
set -u
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -q -m gpu -p no:logging > gpurun_out/r02s_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02s_tests.log
for rep in 1 2; do timeout -k 10 300 python tools/train_probe.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c190-330; done | tee gpurun_out/r02s_train.txt
echo "== actors"; timeout -k 10 300 python tools/train_probe.py --steps 10 --warmup 3 --actors 16 2>&1 | tail -1 | cut -c150-330 | tee -a gpurun_out/r02s_train.txt
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02s_train_launches.csv \
  python tools/train_probe.py --steps 1 --warmup 1 > gpurun_out/r02s_ncu.log 2>&1
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:nff_s -s 4 -c 2 -o gpurun_out/r02s_prof_render \
  env IMAGE_WIDTH=640 python tools/perf_probe.py 0 1 >> gpurun_out/r02s_ncu.log 2>&1
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'nff_|raygen|lidar_decode|dec_' -c 400 --csv \
  --log-file gpurun_out/r02s_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras --no-train --no-decoder --cpu-sample 0 \
  > gpurun_out/r02s_bench_under_ncu.json 2>> gpurun_out/r02s_ncu.log
ls -la gpurun_out/r02s*.ncu-rep
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02s_bench_n1.json 2> gpurun_out/r02s_bench_n1.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/r02s_bench_n1.json
timeout -k 10 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02s_bench_ref.json 2>> gpurun_out/r02s_bench_n1.err; cut -c1-200 gpurun_out/r02s_bench_ref.json
