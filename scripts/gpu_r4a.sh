#!/bin/bash
# kernel (c) bucketed path: parity tests, then single-pass vs bucketed on 50 M / 5 M entries, both occupancies, launch list
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_hubness.py -x -q > $O/r4a_hubness.log 2>&1; echo "pytest rc=$?" >> $O/r4a_hubness.log
tail -5 $O/r4a_hubness.log
for occ in 3 2; do
  echo "== TVC_KOCC_PART_OCC=$occ" >> $O/r4a_probe.log
  TVC_KOCC_PART_OCC=$occ timeout 300 python scripts/probe_kocc.py 50 1000000 >> $O/r4a_probe.log 2>&1
done
timeout 300 python scripts/probe_kocc.py 5 1000000 >> $O/r4a_probe.log 2>&1
timeout 300 python scripts/probe_kocc.py 50 3000000 >> $O/r4a_probe.log 2>&1
timeout 300 python scripts/probe_kocc.py 1 1000000 >> $O/r4a_probe.log 2>&1
cat $O/r4a_probe.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_occurrence -c 60 --csv --log-file $O/r4a_kocc_launches.csv python scripts/probe_kocc.py 50 1000000 > $O/r4a_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r4a_kocc_launches.csv')) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]; ki, vi = h.index('Kernel Name'), h.index('Metric Value')
agg = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    try: agg[r[ki][:60]].append(float(r[vi].replace(',', '')))
    except ValueError: pass
for k, v in agg.items(): print(k, len(v), 'launches, median', sorted(v)[len(v)//2] / 1e3, 'us')
PY
