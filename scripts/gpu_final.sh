#!/bin/bash
# final validation of the round-2 tree: whole GPU suite, smoke(), default bench
O=gpurun_out; mkdir -p $O
timeout 150 python -m pytest tests -m gpu -x -q > $O/r4x_gpu_suite.log 2>&1; echo "pytest rc=$?" >> $O/r4x_gpu_suite.log
tail -3 $O/r4x_gpu_suite.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > $O/r4x_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r4x_smoke.log; tail -3 $O/r4x_smoke.log
timeout 110 python bench.py --no-cpu-baseline > $O/r4x_bench_n1.log 2>&1; echo "bench rc=$?" >> $O/r4x_bench_n1.log
python - <<'PY'
import json
for l in open('gpurun_out/r4x_bench_n1.log'):
    if l.startswith('{'):
        j=json.loads(l)
        print(json.dumps({k:j[k] for k in ('value','ms_per_step','gpu_launches')}), 'e2e', j['e2e']['value'])
        print('roofline', j['roofline']['achieved'], j['roofline']['frac'], 'digest', j['parity']['digest'], j['parity']['ok'])
        print('b', {k:(round(v['us'],1), round(v['frac'],3)) for k,v in j['roofline_b'].items()}, 'c', {k:(round(v['us'],1), round(v['frac'],3)) for k,v in j['roofline_c'].items()})
        print(j['clocks'])
PY
tail -2 $O/r4x_bench_n1.log | cut -c1-300
