#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_apply.py tests/test_partition.py tests/test_gpu_vector.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/d_pytest.log
cat gpurun_out/d_pytest.log
C="3,4,6,f64,9;3,4,6,f64,0;3,4,6,f32,0;3,3,6,f64,0;3,5,5,f64,0;3,2,7,f64,0;3,1,8,f64,0;3,4,5,f64,0;3,4,7,f64,0"
timeout 900 python tools/sweep.py --steps 50 --custom "$C" > gpurun_out/d_sweep.jsonl 2> gpurun_out/d_sweep.err
python - <<'PY'
import json
for l in open('gpurun_out/d_sweep.jsonl'):
    d = json.loads(l)
    print(d.get('p'), d.get('r'), d.get('dtype'), 'req', d.get('requested'), 'variant', d.get('variant'), 'ms %.4f' % d.get('ms', 0), 'gdofs %.1f' % d.get('gdofs', 0), 'frac %.3f' % d.get('roofline_frac', 0), d.get('error', ''))
PY
tail -3 gpurun_out/d_sweep.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-cg 2>&1 | tail -2 | cut -c1-1500
