#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_solver.py -m gpu -x -q 2>&1 | tail -3
python tools/cg_timing.py 300
MFG_CG_UNFUSED=1 python tools/cg_timing.py 300
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 40 --csv --log-file gpurun_out/f_cg_fused.csv python tools/cg_timing.py 60 > /dev/null 2>&1
MFG_CG_UNFUSED=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 40 --csv --log-file gpurun_out/f_cg_unfused.csv python tools/cg_timing.py 60 > /dev/null 2>&1
python - <<'PY'
import csv
for f in ('gpurun_out/f_cg_fused.csv', 'gpurun_out/f_cg_unfused.csv'):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10 and r[0].isdigit()]
    agg = {}
    for r in rows:
        name = r[4][:60]; agg.setdefault(name, []).append(float(r[-1]))
    print(f)
    for k, v in agg.items(): print("   %-62s n=%d mean %.1f us" % (k, len(v), sum(v)/len(v)/ (1000 if max(v) > 5000 else 1)))
PY
