#!/bin/bash
# round-2 GPU session w: whole GPU suite + default bench on the tree with lean issue loops and the resident-query pair kernel
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2w_gpu_suite.log 2>&1; echo "pytest rc=$?" >> $O/r2w_gpu_suite.log
tail -3 $O/r2w_gpu_suite.log
timeout 600 python bench.py > $O/r2w_bench_n1.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2w_bench_n1.log'):
    if l.startswith('{'):
        j=json.loads(l)
        print(json.dumps({k:j[k] for k in ('value','ms_per_step','gpu_launches')}), j['e2e']['value'], j['e2e']['pageable']['value'])
        print(json.dumps(j['roofline'])[:600]); print(json.dumps(j['parity'])[:400]); print(j['clocks'])
PY
