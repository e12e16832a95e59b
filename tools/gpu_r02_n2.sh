#!/usr/bin/env bash
# Round 2, GPU session N2: ncu --set full of the training step's heavy kernels, final binary.
set -u
mkdir -p gpurun_out
timeout -k 10 500 ncu --set full --clock-control none --import-source on -k regex:'mlp_tc_kernel|linear_wgrad_kernel|neurad_encoding_bwd|neurad_encoding_fwd' -s 18 -c 18 -o gpurun_out/r02n2_prof_train \
  python tools/train_probe.py --steps 1 --warmup 1 > gpurun_out/r02n2_ncu.log 2>&1
ls -la gpurun_out/r02n2*.ncu-rep
