#!/bin/bash
# round-2 GPU session j: resident-query kernel - parity vs the pair kernel, then sustained rate
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_search.py -x -q -k "resident_query or pair_kernel" > $O/r2j_tests.log 2>&1; echo "rc=$?" >> $O/r2j_tests.log
tail -15 $O/r2j_tests.log
timeout 400 python scripts/perf_probe2.py 81920x1000000x768 ts_min_tiles=4611686018427387904 ts_min_tiles=48 ts_min_tiles=4611686018427387904 ts_min_tiles=48 > $O/r2j_probe.log 2>&1
cat $O/r2j_probe.log
