#!/usr/bin/env bash
# Round 2, GPU session P: 4x4-blocked wgrad, training MLP forward that keeps the hidden pre-activations -- full GPU suite,
# training step, launch list.
set -u
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -q -m gpu -p no:logging > gpurun_out/r02p_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02p_tests.log
for rep in 1 2; do
  timeout -k 10 300 python tools/train_probe.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c190-330
done | tee gpurun_out/r02p_train_ab.txt
echo "== actors"; timeout -k 10 300 python tools/train_probe.py --steps 10 --warmup 3 --actors 16 2>&1 | tail -1 | cut -c150-330 | tee -a gpurun_out/r02p_train_ab.txt
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02p_train_launches.csv \
  python tools/train_probe.py --steps 1 --warmup 1 > gpurun_out/r02p_ncu.log 2>&1
