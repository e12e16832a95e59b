#!/usr/bin/env bash
# Round 2, GPU session D: A/B of the work-distribution / interpolation / mbarrier variants, full batch and one image.
set -u
mkdir -p gpurun_out
V=neurad-studio_b200/lib/variants
: > gpurun_out/r02d_ab.txt
for lib in neurad-studio_b200/lib/libb200nerf.so $V/libb200nerf_patch.so $V/libb200nerf_f4w.so $V/libb200nerf_f4w_hint.so; do
  echo "== $lib" | tee -a gpurun_out/r02d_ab.txt
  IMAGE_WIDTH=640 NFF_LIB=$lib python tools/perf_probe.py 0 10 2>&1 | tail -1 | tee -a gpurun_out/r02d_ab.txt
  ONE_IMAGE=1 IMAGE_WIDTH=640 NFF_LIB=$lib python tools/perf_probe.py 0 20 2>&1 | tail -1 | tee -a gpurun_out/r02d_ab.txt
done
# correctness of the variants: the fused-render parity tests against each library
for lib in $V/libb200nerf_patch.so $V/libb200nerf_f4w_hint.so; do
  echo "== tests with $lib" | tee -a gpurun_out/r02d_ab.txt
  B200NERF_LIB=$lib python -m pytest tests/test_parity_gpu.py -q -x 2>&1 | tail -3 | tee -a gpurun_out/r02d_ab.txt
done
