#!/usr/bin/env bash
# Round 2, GPU session H: full GPU suite on the current library (tcnn layout, param stream, hybrid work distribution, API
# entry points), then the encoding-backward thread mapping A/B (training step + gradient goldens on the variant).
set -u
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q > gpurun_out/r02h_tests.log 2>&1
echo "tests rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r02h_tests.log | tail -20
V=neurad-studio_b200/lib/variants
for lib in neurad-studio_b200/lib/libb200nerf.so $V/libb200nerf_bwdmap.so; do
  echo "== train step with $lib"
  B200NERF_LIB=$lib timeout -k 10 300 python tools/train_probe.py --steps 5 2>&1 | tail -1 | cut -c150-420
done | tee gpurun_out/r02h_train_ab.txt
B200NERF_LIB=$V/libb200nerf_bwdmap.so timeout -k 10 600 python -m pytest tests/test_zz_module_seams_gpu.py -q -k "gradients or backward or trunc_exp or training" 2>&1 | tail -4 | tee -a gpurun_out/r02h_train_ab.txt
