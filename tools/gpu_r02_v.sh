#!/usr/bin/env bash
# Round 2, GPU session V (N GPUs; final tree: decoder stream in the API arm): multi-GPU gather test, then the bench as the driver launches it.
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02v_topo_n$N.txt 2>&1
timeout -k 10 300 python -m pytest tests/test_multi_gpu.py -q > gpurun_out/r02v_mgpu_test_n$N.log 2>&1; tail -3 gpurun_out/r02v_mgpu_test_n$N.log
timeout -k 10 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 \
  > gpurun_out/r02v_bench_n$N.json 2> gpurun_out/r02v_bench_n$N.err
echo "bench rc=$?"; tail -4 gpurun_out/r02v_bench_n$N.err | cut -c1-300
python - <<PY
import json
d=json.load(open("gpurun_out/r02v_bench_n$N.json"))
print({k:d[k] for k in ("value","ms_per_step","n_gpus","gather_verified","numa")}, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["host_buffers_verified"], d.get("config5_strong",{}).get("value"))
PY
