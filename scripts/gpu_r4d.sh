#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_hubness.py -x -q > $O/r4d_hubness.log 2>&1; echo "pytest rc=$?" >> $O/r4d_hubness.log
tail -3 $O/r4d_hubness.log
for cfg in "2 2" "3 2"; do
  set -- $cfg
  echo "== TVC_KOCC_PART_OCC=$1 TVC_KOCC_PART_GRID=$2" >> $O/r4d_probe.log
  TVC_KOCC_PART_OCC=$1 TVC_KOCC_PART_GRID=$2 timeout 300 python scripts/probe_kocc.py 50 1000000 2>&1 | grep bucketed >> $O/r4d_probe.log
done
timeout 300 python scripts/probe_kocc.py 5 1000000 2>&1 | grep bucketed >> $O/r4d_probe.log
timeout 300 python scripts/probe_kocc.py 50 3000000 2>&1 | grep bucketed >> $O/r4d_probe.log
cat $O/r4d_probe.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_occurrence_(partition|bucket)" -c 60 --csv --log-file $O/r4d_kocc_launches.csv python scripts/probe_kocc.py 50 1000000 > $O/r4d_ncu.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r4d_kocc_launches.csv')) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]; ki, vi = h.index('Kernel Name'), h.index('Metric Value')
print(' '.join(f"{r[ki].split('k_occurrence_')[1][:4]}={float(r[vi].replace(',',''))/1e3:.0f}" for r in rows[hdr+1:]))
PY
