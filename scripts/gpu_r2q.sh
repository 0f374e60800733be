#!/bin/bash
O=gpurun_out; mkdir -p $O
NEVER=4611686018427387904
timeout 200 python -m pytest tests/test_gpu_search.py -x -q -k "resident_query" > $O/r2q_tests.log 2>&1; tail -2 $O/r2q_tests.log
for shape in 81920x1000000x448 81920x1000000x768; do
  echo "== $shape" >> $O/r2q_probe.log
  timeout 300 python scripts/perf_probe2.py $shape rq_min_tiles=$NEVER rq_min_tiles=64 rq_min_tiles=$NEVER,debug_flags=4 >> $O/r2q_probe.log 2>&1
done
cat $O/r2q_probe.log
