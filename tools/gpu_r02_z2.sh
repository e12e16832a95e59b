#!/usr/bin/env bash
# Round 2, GPU session Z2: the full N = 1 bench line with the single-stream API arm (default), bound to the GPU's NUMA node.
set -u
mkdir -p gpurun_out
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02z2_bench_n1.json 2> gpurun_out/r02z2_bench_n1.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02z2_bench_n1.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["traffic"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["host_buffers_verified"], d["e2e"]["decoder_stream"], d["train_step"]["rays_per_s"], d["cpu_baseline"]["value"], d["numa"], d["clocks"])
PY
timeout -k 10 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02z2_bench_n1_b.json 2>> gpurun_out/r02z2_bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02z2_bench_n1_b.json").read().strip().splitlines()[-1])
print("second run:", d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["train_step"]["rays_per_s"])
PY
