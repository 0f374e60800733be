#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python scripts/perf_probe2.py 2048x1000000x768 default pair_min_rows=0 > $O/r3f_probe.log 2>&1
timeout 300 python scripts/perf_probe2.py 3000x50000x768 default pair_min_rows=0 >> $O/r3f_probe.log 2>&1
cat $O/r3f_probe.log
