#!/usr/bin/env bash
# Round 2, GPU session K: run-length aggregated encoding backward -- gradient tests, A/B of the training step across the
# aggregation depths (lib/variants), launch list + ncu of the backward kernels.
set -u
mkdir -p gpurun_out
V=neurad-studio_b200/lib/variants
timeout -k 10 600 python -m pytest tests/test_zz_module_seams_gpu.py -x -q -m gpu -p no:logging > gpurun_out/r02k_seams.log 2>&1; echo "seams rc=$?"; tail -3 gpurun_out/r02k_seams.log
for rep in 1 2; do
  for v in bwd1 agg0 agglo agg62 agghi; do
    echo "== $v"; B200NERF_LIB=$V/libb200nerf_$v.so timeout -k 10 300 python tools/train_probe.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c150-330
  done
  echo "== main (F1=4,F4=2)"; timeout -k 10 300 python tools/train_probe.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c150-330
done | tee gpurun_out/r02k_train_ab.txt
echo "== actors"; timeout -k 10 300 python tools/train_probe.py --steps 10 --warmup 3 --actors 16 2>&1 | tail -1 | cut -c150-330 | tee -a gpurun_out/r02k_train_ab.txt
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02k_train_launches.csv \
  python tools/train_probe.py --steps 1 --warmup 1 > gpurun_out/r02k_ncu.log 2>&1
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:neurad_encoding_bwd -s 4 -c 4 -o gpurun_out/r02k_prof_encoding_bwd \
  python tools/train_probe.py --steps 1 --warmup 1 >> gpurun_out/r02k_ncu.log 2>&1
ls -la gpurun_out/r02k*.ncu-rep
