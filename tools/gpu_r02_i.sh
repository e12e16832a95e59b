#!/usr/bin/env bash
# Round 2, GPU session I: evidence for the benched binary -- full ncu capture of the two render kernels, the launch list of
# bench.py's own command, an ncu capture of the encoding backward (training), and the N = 1 bench line.
set -u
mkdir -p gpurun_out
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:nff_s -s 4 -c 2 -o gpurun_out/r02i_prof_render \
  env IMAGE_WIDTH=640 python tools/perf_probe.py 0 1 > gpurun_out/r02i_ncu.log 2>&1
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'nff_|raygen|lidar_decode|dec_' -c 400 --csv \
  --log-file gpurun_out/r02i_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras --no-train --no-decoder --cpu-sample 0 \
  > gpurun_out/r02i_bench_under_ncu.json 2>> gpurun_out/r02i_ncu.log
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:neurad_encoding_bwd -s 3 -c 3 -o gpurun_out/r02i_prof_encoding_bwd \
  python tools/train_probe.py --steps 1 --warmup 1 >> gpurun_out/r02i_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02i_bench_n1.json 2> gpurun_out/r02i_bench_n1.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/r02i_bench_n1.json
timeout -k 10 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02i_bench_ref.json 2>> gpurun_out/r02i_bench_n1.err; cut -c600-900 gpurun_out/r02i_bench_ref.json
