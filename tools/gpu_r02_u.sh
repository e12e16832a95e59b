#!/usr/bin/env bash
# Round 2, GPU session U: compute-sanitizer (memcheck, racecheck) over one small training step and the MLP operator tests --
# the kernels that gained cp.async pipelines / shared-memory staging this round.
set -u
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  echo "== $tool: training step (2048 + 1024 rays, small tables)"
  timeout -k 10 420 $CS --tool $tool --print-limit 20 python tools/train_probe.py --cam-rays 2048 --lidar-rays 1024 --steps 1 --warmup 0 --small-tables \
    > gpurun_out/r02u_${tool}_train.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error:|hazard" gpurun_out/r02u_${tool}_train.log | head -8
  echo "== $tool: training step with 16 actors"
  timeout -k 10 420 $CS --tool $tool --print-limit 20 python tools/train_probe.py --cam-rays 1024 --lidar-rays 512 --steps 1 --warmup 0 --small-tables --actors 16 \
    > gpurun_out/r02u_${tool}_train_actors.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error:|hazard" gpurun_out/r02u_${tool}_train_actors.log | head -8
done
echo "== memcheck: MLP operator tests (1000 rows)"
timeout -k 10 600 $CS --tool memcheck --print-limit 20 python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:logging -k "mlp_fwd and 1000" \
  > gpurun_out/r02u_memcheck_mlp.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02u_memcheck_mlp.log | head
