#!/usr/bin/env bash
# Round 2, GPU session L: decoder gradient from saved features; A/B of aggregation depth / register budget / segment count.
set -u
mkdir -p gpurun_out
V=neurad-studio_b200/lib/variants
timeout -k 10 600 python -m pytest tests/test_zz_module_seams_gpu.py -x -q -m gpu -p no:logging > gpurun_out/r02l_seams.log 2>&1; echo "seams rc=$?"; tail -3 gpurun_out/r02l_seams.log
for rep in 1 2; do
  for v in a62_4 a63_4 a63_2 a64_2 a65_2 a64_2_s4 a64_2_s16; do
    echo "== $v"; B200NERF_LIB=$V/libb200nerf_$v.so timeout -k 10 300 python tools/train_probe.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c190-330
  done
  echo "== main (F1=4,F4=2)"; timeout -k 10 300 python tools/train_probe.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c190-330
done | tee gpurun_out/r02l_train_ab.txt
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02l_train_launches.csv \
  env B200NERF_LIB=$V/libb200nerf_a64_2.so python tools/train_probe.py --steps 1 --warmup 1 > gpurun_out/r02l_ncu.log 2>&1
