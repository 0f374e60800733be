#!/bin/bash
# round-2 GPU session k: how thin may the pair kernel's ring be?  (debug bits 4-6 = stages used; 16*n)
O=gpurun_out; mkdir -p $O
timeout 600 python scripts/perf_probe2.py 81920x1000000x768 default debug_flags=80 debug_flags=64 debug_flags=48 debug_flags=32 default debug_flags=64 > $O/r2k_ring.log 2>&1
cat $O/r2k_ring.log
