#!/bin/bash
O=gpurun_out; mkdir -p $O
NEVER=4611686018427387904
timeout 300 python -m pytest tests/test_gpu_search.py -x -q -k "resident_query or pair_kernel" > $O/r2v_tests.log 2>&1; tail -3 $O/r2v_tests.log
for shape in 81920x1000000x768 81920x1000000x512; do
  echo "== $shape" >> $O/r2v_probe.log
  timeout 300 python scripts/perf_probe2.py $shape rq_min_tiles=$NEVER rq_min_tiles=64,rq_resident=5 rq_min_tiles=64,rq_resident=6 rq_min_tiles=64,rq_resident=7 rq_min_tiles=$NEVER rq_min_tiles=64,rq_resident=5 >> $O/r2v_probe.log 2>&1
done
cat $O/r2v_probe.log
