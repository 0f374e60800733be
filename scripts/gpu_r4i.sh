#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_hubness.py -x -q > $O/r4i_hubness.log 2>&1; echo "pytest rc=$?" >> $O/r4i_hubness.log
TVC_KOCC_PART_KIND=1 timeout 600 python -m pytest tests/test_gpu_hubness.py -x -q -k bucketed >> $O/r4i_hubness.log 2>&1; echo "pytest (general kernel) rc=$?" >> $O/r4i_hubness.log
grep -E "passed|failed|rc=" $O/r4i_hubness.log
for cfg in "3 0" "2 0" "3 1"; do
  set -- $cfg
  echo "== TVC_KOCC_PART_OCC=$1 TVC_KOCC_PART_KIND=$2" >> $O/r4i_probe.log
  TVC_KOCC_PART_OCC=$1 TVC_KOCC_PART_KIND=$2 timeout 300 python scripts/probe_kocc.py 50 1000000 2>&1 | grep bucketed >> $O/r4i_probe.log
done
timeout 300 python scripts/probe_kocc.py 5 1000000 2>&1 | grep bucketed >> $O/r4i_probe.log
timeout 300 python scripts/probe_kocc.py 50 3000000 2>&1 | grep bucketed >> $O/r4i_probe.log
cat $O/r4i_probe.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_occurrence_(partition|bucket)" -c 60 --csv --log-file $O/r4i_kocc_launches.csv python scripts/probe_kocc.py 50 1000000 > $O/r4i_ncu.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r4i_kocc_launches.csv')) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]; ki, vi = h.index('Kernel Name'), h.index('Metric Value')
print(' '.join(f"{r[ki].split('k_occurrence_')[1][:4]}={float(r[vi].replace(',',''))/1e3:.0f}" for r in rows[hdr+1:]))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_occurrence_partition_tma" --launch-skip 3 -c 1 -o $O/r4i_partition python scripts/probe_kocc.py 50 1000000 > $O/r4i_ncu2.log 2>&1
