#!/bin/bash
# round-2 GPU session m: resident-query pair kernel at d = 512 / 448 / 384 (everything or nearly everything resident)
O=gpurun_out; mkdir -p $O
NEVER=4611686018427387904
for shape in 81920x1000000x512 81920x1000000x448 81920x1000000x640; do
  echo "== $shape" >> $O/r2m_probe.log
  timeout 300 python scripts/perf_probe2.py $shape rq_min_tiles=$NEVER rq_min_tiles=64 rq_min_tiles=$NEVER rq_min_tiles=64 >> $O/r2m_probe.log 2>&1
done
cat $O/r2m_probe.log
