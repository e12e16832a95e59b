#!/usr/bin/env bash
# Round 2, GPU session O: cp.async double-buffered MLP operator, register-resident forward encoding with coalesced stores,
# warp-merged pending sums + input prefetch in the encoding backward -- full GPU suite, training-step A/B, launch list.
set -u
mkdir -p gpurun_out
V=neurad-studio_b200/lib/variants
timeout -k 10 900 python -m pytest tests -x -q -m gpu -p no:logging > gpurun_out/r02o_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02o_tests.log
for rep in 1 2; do
  for v in wm0 wm6 nopf; do
    echo "== $v"; B200NERF_LIB=$V/libb200nerf_$v.so timeout -k 10 300 python tools/train_probe.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c190-330
  done
  echo "== main (wm3, prefetch)"; timeout -k 10 300 python tools/train_probe.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c190-330
done | tee gpurun_out/r02o_train_ab.txt
echo "== actors"; timeout -k 10 300 python tools/train_probe.py --steps 10 --warmup 3 --actors 16 2>&1 | tail -1 | cut -c150-330 | tee -a gpurun_out/r02o_train_ab.txt
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02o_train_launches.csv \
  python tools/train_probe.py --steps 1 --warmup 1 > gpurun_out/r02o_ncu.log 2>&1
