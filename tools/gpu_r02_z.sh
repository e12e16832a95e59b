#!/usr/bin/env bash
# Round 2, GPU session Z: e2e repeatability at N = 1 with the process bound to the GPU's NUMA node, then the full bench line.
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02z_topo.txt 2>&1
for rep in 1 2 3; do
  timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-extras --no-train --no-decoder --cpu-sample 0 2> gpurun_out/r02z_err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print(round(d['value']/1e6,2),'M value;', 'e2e', round(e['value']/1e6,2),'M rays/s', round(e['ms_per_step'],2),'ms', d['numa'])"
done | tee gpurun_out/r02z_e2e_repeat.txt
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02z_bench_n1.json 2> gpurun_out/r02z_bench_n1.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02z_bench_n1.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["traffic"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["host_buffers_verified"], d["train_step"]["rays_per_s"], d["cpu_baseline"]["value"], d["numa"], d["clocks"])
PY
