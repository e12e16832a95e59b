#!/usr/bin/env bash
# Round 2, GPU session G: hybrid work distribution (full rounds grid-stride + balanced remainder) and mbarrier hints, A/B;
# fused-render parity tests on the new default; then the training-step profile (session F).
set -u
mkdir -p gpurun_out
V=neurad-studio_b200/lib/variants
: > gpurun_out/r02g_ab.txt
for lib in neurad-studio_b200/lib/libb200nerf.so $V/libb200nerf_hint2k.so $V/libb200nerf_hint20k.so; do
  echo "== $lib" | tee -a gpurun_out/r02g_ab.txt
  IMAGE_WIDTH=640 NFF_LIB=$lib python tools/perf_probe.py 0 10 2>&1 | tail -1 | cut -c1-60 | tee -a gpurun_out/r02g_ab.txt
  ONE_IMAGE=1 IMAGE_WIDTH=640 NFF_LIB=$lib python tools/perf_probe.py 0 20 2>&1 | tail -1 | cut -c1-60 | tee -a gpurun_out/r02g_ab.txt
done
python -m pytest tests/test_parity_gpu.py -q 2>&1 | tail -3 | tee -a gpurun_out/r02g_ab.txt
bash tools/gpu_r02_f.sh
