#!/bin/bash
O=gpurun_out; mkdir -p $O
NEVER=4611686018427387904
timeout 300 python scripts/perf_probe2.py 81920x1000000x448 rq_min_tiles=$NEVER rq_min_tiles=$NEVER,debug_flags=4 rq_min_tiles=64 rq_min_tiles=$NEVER,debug_flags=68 rq_min_tiles=$NEVER,debug_flags=64 >> $O/r2p_probe.log 2>&1
cat $O/r2p_probe.log
