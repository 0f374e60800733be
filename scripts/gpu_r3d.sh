#!/bin/bash
O=gpurun_out; mkdir -p $O
NEVER=4611686018427387904
timeout 300 python -m pytest tests/test_gpu_search.py -x -q > $O/r3d_tests.log 2>&1; tail -2 $O/r3d_tests.log
for shape in 81920x1000000x768; do
  echo "== $shape" >> $O/r3d_probe.log
  timeout 300 python scripts/perf_probe2.py $shape rq_min_tiles=64,rq_resident=8 rq_min_tiles=64,rq_resident=9 rq_min_tiles=64,rq_resident=8 rq_min_tiles=64,rq_resident=9 >> $O/r3d_probe.log 2>&1
done
cat $O/r3d_probe.log
