#!/usr/bin/env bash
# Round 2, GPU session F: where the training step's time goes (per-kernel device times of one step), wgrad A/B.
set -u
mkdir -p gpurun_out
python tools/train_probe.py --steps 5 > gpurun_out/r02f_train.json 2> gpurun_out/r02f_train.err; cat gpurun_out/r02f_train.json
B200NERF_WGRAD=tc python tools/train_probe.py --steps 5 > gpurun_out/r02f_train_wgrad_tc.json 2>> gpurun_out/r02f_train.err; cat gpurun_out/r02f_train_wgrad_tc.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02f_train_launches.csv \
    python tools/train_probe.py --steps 1 --warmup 1 > gpurun_out/r02f_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r02f_train_launches.csv")) if len(r)>10 and r[0].isdigit()]
# second half = the timed step (warm-up step first)
half=rows[len(rows)//2:]
tot=collections.Counter(); cnt=collections.Counter()
for r in half:
    name=r[4].split("(")[0][:70]; tot[name]+=float(r[-1]); cnt[name]+=1
s=sum(tot.values())
print(f"{len(half)} launches, {s/1e6:.2f} ms of kernel time in the step")
for k,v in tot.most_common(25): print(f"{v/1e6:8.3f} ms  {100*v/s:5.1f}%  x{cnt[k]:<4d} {k}")
PY
