#!/usr/bin/env bash
# Round 2, GPU session N2: ncu --set full of the training step's heavy kernels, final binary.  The report (> 64 MiB with the
# sources imported) stays on the box; only its raw page comes back as CSV.
set -u
mkdir -p gpurun_out
timeout -k 10 500 ncu --set full --clock-control none -k regex:'mlp_tc_kernel|linear_wgrad_kernel|neurad_encoding_bwd|neurad_encoding_fwd' -s 18 -c 18 -o /tmp/r02n2_prof_train \
  python tools/train_probe.py --steps 1 --warmup 1 > gpurun_out/r02n2_ncu.log 2>&1
ncu -i /tmp/r02n2_prof_train.ncu-rep --page raw --csv > gpurun_out/r02n2_train_raw.csv 2>/dev/null
ls -la /tmp/r02n2_prof_train.ncu-rep gpurun_out/r02n2_train_raw.csv
