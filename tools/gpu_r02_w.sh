#!/usr/bin/env bash
# Round 2, GPU session W: the N = 1 bench line of the committed tree (traffic.json now matches the sources), smoke(), and
# compute-sanitizer over a small fused render + the rgb decoder.
set -u
mkdir -p gpurun_out
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02w_bench_n1.json 2> gpurun_out/r02w_bench_n1.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02w_bench_n1.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["traffic"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["host_buffers_verified"], d["train_step"]["rays_per_s"], d["clocks"])
PY
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  timeout -k 10 420 $CS --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02w_${tool}_smoke.log 2>&1; echo "$tool smoke rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r02w_${tool}_smoke.log
done
timeout -k 10 600 $CS --tool memcheck --print-limit 20 python -m pytest tests/test_zz_module_seams_gpu.py -q -m gpu -p no:logging -k "metric_entry_points or get_outputs" \
  > gpurun_out/r02w_memcheck_api.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02w_memcheck_api.log | head
