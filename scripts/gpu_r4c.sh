#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_occurrence_partition" --launch-skip 3 -c 1 -o $O/r4e_partition python scripts/probe_kocc.py 50 1000000 > $O/r4e_ncu.log 2>&1
tail -3 $O/r4e_ncu.log
ls -la $O/r4e_partition.ncu-rep
