#!/bin/bash
# pass 2 with L lanes per segment and a visit-weighted deal: parity, then the 50 M stream (every step under a short timeout)
O=gpurun_out; mkdir -p $O
timeout 60 python -m pytest tests/test_gpu_hubness.py -x -q > $O/r4k_hubness.log 2>&1; echo "pytest rc=$?" >> $O/r4k_hubness.log
TVC_KOCC_PART_KIND=1 TVC_KOCC_PART_LANES=3 TVC_KOCC_PART_VISIT=200 timeout 60 python -m pytest tests/test_gpu_hubness.py -x -q -k bucketed >> $O/r4k_hubness.log 2>&1; echo "pytest (general kernel, 8 lanes, visit 200) rc=$?" >> $O/r4k_hubness.log
grep -E "passed|failed|rc=" $O/r4k_hubness.log
for cfg in "0 0" "5 0" "0 128" "0 512"; do
  set -- $cfg
  echo "== TVC_KOCC_PART_LANES=$1 TVC_KOCC_PART_VISIT=$2" >> $O/r4k_probe.log
  TVC_KOCC_PART_LANES=$1 TVC_KOCC_PART_VISIT=$2 timeout 40 python scripts/probe_kocc.py 50 1000000 2>&1 | grep bucketed >> $O/r4k_probe.log
done
cat $O/r4k_probe.log
