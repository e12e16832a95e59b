#!/usr/bin/env bash
# Round 2, GPU session J: register-resident encoding backward -- gradient tests, A/B of the training step against the
# previous binary (lib/variants/libb200nerf_bwd1.so), launch list + ncu of the new backward kernels, re-capture of the render
# kernels for profiles/traffic.json (b200nerf.cu changed: the launcher of the backward), full GPU suite, N = 1 bench line.
set -u
mkdir -p gpurun_out
V=neurad-studio_b200/lib/variants
timeout -k 10 600 python -m pytest tests/test_zz_module_seams_gpu.py -x -q -m gpu -p no:logging > gpurun_out/r02j_seams.log 2>&1; echo "seams rc=$?"; tail -3 gpurun_out/r02j_seams.log
for rep in 1 2; do
  echo "== old"; B200NERF_LIB=$V/libb200nerf_bwd1.so timeout -k 10 300 python tools/train_probe.py --steps 10 --warmup 3 2>&1 | tail -2
  echo "== new"; timeout -k 10 300 python tools/train_probe.py --steps 10 --warmup 3 2>&1 | tail -2
done | tee gpurun_out/r02j_train_ab.txt
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02j_train_launches.csv \
  python tools/train_probe.py --steps 1 --warmup 1 > gpurun_out/r02j_ncu.log 2>&1
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:neurad_encoding_bwd -s 4 -c 4 -o gpurun_out/r02j_prof_encoding_bwd \
  python tools/train_probe.py --steps 1 --warmup 1 >> gpurun_out/r02j_ncu.log 2>&1
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:nff_s -s 4 -c 2 -o gpurun_out/r02j_prof_render \
  env IMAGE_WIDTH=640 python tools/perf_probe.py 0 1 >> gpurun_out/r02j_ncu.log 2>&1
ls -la gpurun_out/r02j*.ncu-rep
timeout -k 10 900 python -m pytest tests -x -q -m gpu -p no:logging > gpurun_out/r02j_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02j_tests.log
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02j_bench_n1.json 2> gpurun_out/r02j_bench_n1.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/r02j_bench_n1.json
