#!/usr/bin/env bash
# Round 2, GPU session T: API arm with the rgb decoder on a side stream (pipelined under the next image's render) -- GPU tests
# of the entry points, A/B of bench.py's e2e with and without it.
set -u
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_zz_module_seams_gpu.py -q -m gpu -p no:logging > gpurun_out/r02t_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02t_tests.log
for rep in 1 2; do
  for ds in 0 1; do
    echo "== decoder_stream=$ds"; B200_E2E_DECODER_STREAM=$ds timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-extras --no-train --no-decoder --cpu-sample 0 2> gpurun_out/r02t_err_$ds.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print(round(d['value']/1e6,2),'M value;', 'e2e', round(e['value']/1e6,2),'M rays/s', round(e['ms_per_step'],2),'ms verified', e['host_buffers_verified'], e.get('decoder_stream'))"
  done
done | tee gpurun_out/r02t_e2e_ab.txt
