#!/usr/bin/env bash
# First GPU session of the next round: everything that was written after round 1's GPU budget ran out, in the order that
# gives verdicts soonest.  Run each block as ONE gpurun call (outputs land in gpurun_out/), e.g.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/next_gpu_session.sh tests'
set -u
mkdir -p gpurun_out
case "${1:-tests}" in
  tests)
    # 1. the GPU tests that have never run (module seams added late, compat tests, experimental tcgen05 wgrad), then the rest
    python -m pytest tests/test_zz_module_seams_gpu.py tests/test_zz_reference_compat_gpu.py -x -q -rA > gpurun_out/zz_tests.log 2>&1
    python -m pytest tests/test_zzz_experimental_gpu.py -q -rA > gpurun_out/experimental_tests.log 2>&1
    python -m pytest tests -x -q -m gpu > gpurun_out/all_gpu_tests.log 2>&1
    tail -3 gpurun_out/zz_tests.log gpurun_out/experimental_tests.log gpurun_out/all_gpu_tests.log
    ;;
  train)
    # 2. the training step: device times, CUDA-core vs tcgen05 weight gradient, with and without actors
    python tools/train_probe.py --steps 10 > gpurun_out/train_probe.json 2> gpurun_out/train_probe.err
    B200NERF_WGRAD=tc python tools/train_probe.py --steps 10 > gpurun_out/train_probe_wgrad_tc.json 2>> gpurun_out/train_probe.err
    python tools/train_probe.py --steps 10 --actors 16 > gpurun_out/train_probe_actors16.json 2>> gpurun_out/train_probe.err
    cat gpurun_out/train_probe*.json
    ;;
  ncu)
    # 3. where the training step's time goes: launch list, then full captures of the two heaviest backward kernels
    ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/train_launches.csv \
        python tools/train_probe.py --steps 1 --warmup 1 > gpurun_out/train_ncu.log 2>&1
    python tools/ncu_by_function.py gpurun_out/train_launches.csv > gpurun_out/train_launches.txt 2>&1 || true
    ncu --set full --clock-control none --import-source on -k regex:neurad_encoding_bwd -c 2 -o gpurun_out/prof_encoding_bwd \
        python tools/train_probe.py --steps 1 --warmup 1 >> gpurun_out/train_ncu.log 2>&1
    ncu --set full --clock-control none --import-source on -k regex:linear_wgrad -c 2 -o gpurun_out/prof_wgrad \
        python tools/train_probe.py --steps 1 --warmup 1 >> gpurun_out/train_ncu.log 2>&1
    head -30 gpurun_out/train_launches.txt
    ;;
  bench)
    python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
    ;;
  *) echo "usage: $0 tests|train|ncu|bench"; exit 2 ;;
esac
