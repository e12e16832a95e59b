#!/usr/bin/env bash
# Round 2, GPU session C: the restructured bench at N = 1 (all legs) + the GPU tests touched since session B.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_zz_module_seams_gpu.py -q -k "metric_entry or module_walk or trunc_exp" > gpurun_out/r02c_tests.log 2>&1
echo "tests rc=$?"; grep -E "^FAILED|^ERROR|^E  |passed|failed" gpurun_out/r02c_tests.log | tail -12
python bench.py --steps 10 --warmup 3 > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err
echo "bench rc=$?"; tail -5 gpurun_out/r02c_bench_n1.err; cat gpurun_out/r02c_bench_n1.json | cut -c1-6000
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02c_bench_ref.json 2>> gpurun_out/r02c_bench_n1.err; cut -c1-400 gpurun_out/r02c_bench_ref.json
