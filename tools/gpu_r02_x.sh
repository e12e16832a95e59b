#!/usr/bin/env bash
# Round 2, GPU session X: CTA-level shared-memory collection of the coarse cells in the scatter backward -- gradient tests,
# A/B of the training step (none / F1 only / default 3+2 / 6+4 levels), then the items of session W (N = 1 bench line of the
# committed tree, smoke, compute-sanitizer over a small fused render and the API entry points).
set -u
mkdir -p gpurun_out
V=neurad-studio_b200/lib/variants
timeout -k 10 600 python -m pytest tests/test_zz_module_seams_gpu.py -q -m gpu -p no:logging > gpurun_out/r02x_seams.log 2>&1; echo "seams rc=$?"; tail -3 gpurun_out/r02x_seams.log
for rep in 1 2; do
  for v in cta0 cta30 cta64; do
    echo "== $v"; B200NERF_LIB=$V/libb200nerf_$v.so timeout -k 10 300 python tools/train_probe.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c190-330
  done
  echo "== main (3 + 2 levels)"; timeout -k 10 300 python tools/train_probe.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c190-330
done | tee gpurun_out/r02x_train_ab.txt
echo "== actors"; timeout -k 10 300 python tools/train_probe.py --steps 10 --warmup 3 --actors 16 2>&1 | tail -1 | cut -c150-330 | tee -a gpurun_out/r02x_train_ab.txt
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  timeout -k 10 420 $CS --tool $tool --print-limit 20 python tools/train_probe.py --cam-rays 2048 --lidar-rays 1024 --steps 1 --warmup 0 --small-tables --actors 16 \
    > gpurun_out/r02x_${tool}_train.log 2>&1; echo "$tool train rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r02x_${tool}_train.log
done
bash tools/gpu_r02_w.sh
