#!/usr/bin/env bash
# Round 2, GPU session Y (final): full GPU suite and the N = 1 bench line + reference arm of the committed tree.
set -u
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -q -m gpu -p no:logging > gpurun_out/r02y_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02y_tests.log
timeout -k 10 300 python tools/train_probe.py --steps 10 --warmup 3 --actors 16 2>&1 | tail -1 | cut -c150-330
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02y_bench_n1.json 2> gpurun_out/r02y_bench_n1.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02y_bench_n1.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["traffic"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["host_buffers_verified"], d["train_step"]["rays_per_s"], d["clocks"])
PY
timeout -k 10 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02y_bench_ref.json 2>> gpurun_out/r02y_bench_n1.err; cut -c1-120 gpurun_out/r02y_bench_ref.json
