#!/bin/bash
O=gpurun_out; mkdir -p $O
for g in 1 3 4; do
  echo "== TVC_KOCC_PART_GRID=$g" >> $O/r4j_probe.log
  TVC_KOCC_PART_GRID=$g timeout 300 python scripts/probe_kocc.py 50 1000000 2>&1 | grep bucketed >> $O/r4j_probe.log
done
echo "== sizes (default grid)" >> $O/r4j_probe.log
for mm in 1 2 3 10 20; do timeout 300 python scripts/probe_kocc.py $mm 1000000 2>&1 | grep "skewed" >> $O/r4j_probe.log; done
timeout 300 python scripts/probe_kocc.py 50 500000 2>&1 | grep "skewed" >> $O/r4j_probe.log
timeout 300 python scripts/probe_kocc.py 50 118287 2>&1 | grep "skewed" >> $O/r4j_probe.log
cat $O/r4j_probe.log
