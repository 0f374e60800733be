#!/bin/bash
O=gpurun_out; mkdir -p $O
NEVER=4611686018427387904
timeout 300 python scripts/perf_probe2.py 81920x1000000x448 rq_min_tiles=$NEVER rq_min_tiles=64 rq_min_tiles=$NEVER,debug_flags=1 rq_min_tiles=64,debug_flags=1 rq_min_tiles=64,pace_every=0 rq_min_tiles=64,debug_flags=2 >> $O/r2o_probe.log 2>&1
cat $O/r2o_probe.log
