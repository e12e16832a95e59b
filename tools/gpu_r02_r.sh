#!/usr/bin/env bash
# Round 2, GPU session R: batched mask loads in the dgrad stores, 8-byte vector reductions for x-adjacent corners (F = 1) --
# full GPU suite, training step A/B, launch list.
set -u
mkdir -p gpurun_out
V=neurad-studio_b200/lib/variants
timeout -k 10 900 python -m pytest tests -q -m gpu -p no:logging > gpurun_out/r02r_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02r_tests.log
for rep in 1 2; do
  echo "== main"; timeout -k 10 300 python tools/train_probe.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c190-330
  echo "== pair0"; B200NERF_LIB=$V/libb200nerf_pair0.so timeout -k 10 300 python tools/train_probe.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c190-330
done | tee gpurun_out/r02r_train_ab.txt
echo "== actors"; timeout -k 10 300 python tools/train_probe.py --steps 10 --warmup 3 --actors 16 2>&1 | tail -1 | cut -c150-330 | tee -a gpurun_out/r02r_train_ab.txt
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02r_train_launches.csv \
  python tools/train_probe.py --steps 1 --warmup 1 > gpurun_out/r02r_ncu.log 2>&1
