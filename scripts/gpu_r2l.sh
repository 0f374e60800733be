#!/bin/bash
# round-2 GPU session l: pair kernel with a resident query tile - parity, then sustained rate A/B
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_search.py -x -q -k "resident_query or pair_kernel" > $O/r2l_tests.log 2>&1; echo "rc=$?" >> $O/r2l_tests.log
tail -15 $O/r2l_tests.log
NEVER=4611686018427387904
timeout 400 python scripts/perf_probe2.py 81920x1000000x768 rq_min_tiles=$NEVER rq_min_tiles=64 rq_min_tiles=$NEVER rq_min_tiles=64 > $O/r2l_probe.log 2>&1
cat $O/r2l_probe.log
