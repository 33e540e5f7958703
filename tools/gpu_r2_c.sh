#!/bin/bash
mkdir -p gpurun_out
C=""
for v in 9 51 53 55 56 57 58; do C="$C;3,4,6,f64,$v"; done
for v in 9 51 53 55 56 57 58; do C="$C;3,4,6,f32,$v"; done
for v in 9 51 53 56 58; do C="$C;3,3,6,f64,$v"; done
for v in 9 51 53; do C="$C;3,5,5,f64,$v"; done
for v in 2 51 53 56 58; do C="$C;3,2,7,f64,$v"; done
timeout 900 python tools/sweep.py --steps 50 --custom "${C:1}" > gpurun_out/c_sweep.jsonl 2> gpurun_out/c_sweep.err
python - <<'PY'
import json
for l in open('gpurun_out/c_sweep.jsonl'):
    d = json.loads(l)
    print(d.get('p'), d.get('r'), d.get('dtype'), 'req', d.get('requested'), 'ms %.4f' % d.get('ms', 0), 'frac %.3f' % d.get('roofline_frac', 0), d.get('error', ''))
PY
tail -3 gpurun_out/c_sweep.err
