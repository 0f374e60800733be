#!/bin/bash
# local-memory slow-path staging: 14 units (resident 6 / 7 / 8 + ring 8 / 7 / 6)
O=gpurun_out; mkdir -p $O
NEVER=4611686018427387904
timeout 300 python -m pytest tests/test_gpu_search.py -x -q -k "resident_query or pair_kernel or ties or skip_self" > $O/r3c_tests.log 2>&1; tail -3 $O/r3c_tests.log
for shape in 81920x1000000x768 81920x1000000x512; do
  echo "== $shape" >> $O/r3c_probe.log
  timeout 300 python scripts/perf_probe2.py $shape rq_min_tiles=$NEVER rq_min_tiles=64,rq_resident=6 rq_min_tiles=64,rq_resident=7 rq_min_tiles=64,rq_resident=8 rq_min_tiles=$NEVER rq_min_tiles=64,rq_resident=7 >> $O/r3c_probe.log 2>&1
done
cat $O/r3c_probe.log
