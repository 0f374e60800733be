#!/bin/bash
O=gpurun_out; mkdir -p $O
NEVER=4611686018427387904
timeout 300 python scripts/perf_probe2.py 81920x1000000x448 rq_min_tiles=$NEVER,debug_flags=4 rq_min_tiles=$NEVER,debug_flags=12 rq_min_tiles=$NEVER,debug_flags=4 rq_min_tiles=$NEVER,debug_flags=12 > $O/r2s_probe.log 2>&1
cat $O/r2s_probe.log
