#!/bin/bash
O=gpurun_out; mkdir -p $O
NEVER=4611686018427387904
timeout 400 python -m pytest tests/test_gpu_search.py -x -q > $O/r2t_tests.log 2>&1; tail -3 $O/r2t_tests.log
timeout 300 python scripts/perf_probe2.py 81920x1000000x768 rq_min_tiles=$NEVER rq_min_tiles=$NEVER,debug_flags=4 rq_min_tiles=$NEVER,debug_flags=1 rq_min_tiles=$NEVER > $O/r2t_probe.log 2>&1
cat $O/r2t_probe.log
