#!/bin/bash
# final validation of the round-2 tree: memcheck of the new GEMM kernels, whole GPU suite, smoke(), default bench, reference arm
O=gpurun_out; mkdir -p $O
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_search.py -x -q -k "resident_query_kernels_match_pair_kernel and (700 or 513 or 640)" > $O/r3e_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/r3e_memcheck.log
tail -4 $O/r3e_memcheck.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/r3e_gpu_suite.log 2>&1; echo "pytest rc=$?" >> $O/r3e_gpu_suite.log
tail -3 $O/r3e_gpu_suite.log
timeout 300 python __graft_entry__.py smoke > $O/r3e_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r3e_smoke.log; tail -3 $O/r3e_smoke.log
timeout 600 python bench.py > $O/r3e_bench_n1.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r3e_bench_n1.log'):
    if l.startswith('{'):
        j=json.loads(l)
        print(json.dumps({k:j[k] for k in ('value','ms_per_step','gpu_launches')}), 'e2e', j['e2e']['value'], 'pageable', j['e2e']['pageable']['value'])
        print('roofline', j['roofline']['achieved'], j['roofline']['frac'], 'digest', j['parity']['digest'], j['parity']['ok'])
        print('b', {k:(round(v['us'],1), round(v['frac'],3)) for k,v in j['roofline_b'].items()}, 'c', {k:(round(v['us'],1), round(v['frac'],3)) for k,v in j['roofline_c'].items()})
        print('latency', j['latency']['single_query_ms'], 'dropin', {k:round(v) for k,v in j['dropin'].items() if isinstance(v,(int,float))})
        print('cpu', j['cpu_baseline']['value'], j['cpu_baseline']['cores'], j['clocks'])
PY
