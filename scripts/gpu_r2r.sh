#!/bin/bash
O=gpurun_out; mkdir -p $O
NEVER=4611686018427387904
for rq in $NEVER 64; do
TVC_RQ_MIN_TILES=$rq timeout 400 ncu --set full --import-source on --clock-control none -k regex:gemm_topk_pair --launch-skip 1 --launch-count 1 -f -o $O/r2r_rq$rq python scripts/perf_probe2.py 81920x1000000x448 default > $O/r2r_ncu_$rq.log 2>&1
done
ls -la $O | grep r2r
