#!/usr/bin/env bash
# Round 2, GPU session A: the full GPU suite on the new addressing + tightened parity asserts, then A/B timings.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r02a_tests.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/r02a_tests.log
V=neurad-studio_b200/lib/variants
for lib in $V/libb200nerf_r01.so neurad-studio_b200/lib/libb200nerf.so $V/libb200nerf_parity.so; do
  echo "== $lib"; IMAGE_WIDTH=640 NFF_LIB=$lib python tools/perf_probe.py 0 10 2>&1 | tail -1
done | tee gpurun_out/r02a_ab.txt
echo "== 16 actors (config 3)"; IMAGE_WIDTH=640 python tools/perf_probe.py 16 10 2>&1 | tail -1 | tee -a gpurun_out/r02a_ab.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02a_launches.csv \
  env IMAGE_WIDTH=640 python tools/perf_probe.py 0 3 > gpurun_out/r02a_ncu.log 2>&1
grep -E "nff_|raygen" gpurun_out/r02a_launches.csv | tail -8
