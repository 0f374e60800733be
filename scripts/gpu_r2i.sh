#!/bin/bash
# round-2 GPU session i: what would a resident query operand save?  (debug bit 4: the A tile is not re-loaded)
O=gpurun_out; mkdir -p $O
timeout 600 python scripts/perf_probe2.py 81920x1000000x768 default debug_flags=4 default debug_flags=4 debug_flags=1 debug_flags=5 > $O/r2i_probe_a.log 2>&1
cat $O/r2i_probe_a.log
