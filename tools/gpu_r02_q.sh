#!/usr/bin/env bash
# Round 2, GPU session Q: cp.async double-buffered wgrad, dgrad with fused ReLU mask -- full GPU suite, training step,
# launch list; timing probes that drop the scatter of the first 1 / 3 / all levels (what do the hot rows cost?).
set -u
mkdir -p gpurun_out
V=neurad-studio_b200/lib/variants
timeout -k 10 900 python -m pytest tests -q -m gpu -p no:logging > gpurun_out/r02q_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02q_tests.log
for rep in 1 2; do
  echo "== main"; timeout -k 10 300 python tools/train_probe.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c190-330
  for v in skip1 skip3 skip8; do
    echo "== probe $v"; B200NERF_LIB=$V/libb200nerf_$v.so timeout -k 10 300 python tools/train_probe.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c190-330
  done
done | tee gpurun_out/r02q_train_ab.txt
echo "== actors"; timeout -k 10 300 python tools/train_probe.py --steps 10 --warmup 3 --actors 16 2>&1 | tail -1 | cut -c150-330 | tee -a gpurun_out/r02q_train_ab.txt
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02q_train_launches.csv \
  python tools/train_probe.py --steps 1 --warmup 1 > gpurun_out/r02q_ncu.log 2>&1
for v in skip3 skip8; do
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:neurad_encoding_bwd -c 16 --csv --log-file gpurun_out/r02q_bwd_$v.csv \
  env B200NERF_LIB=$V/libb200nerf_$v.so python tools/train_probe.py --steps 1 --warmup 1 >> gpurun_out/r02q_ncu.log 2>&1
done
