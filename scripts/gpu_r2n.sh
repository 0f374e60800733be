#!/bin/bash
O=gpurun_out; mkdir -p $O
NEVER=4611686018427387904
for shape in 81920x1000000x448 81920x1000000x768; do
  echo "== $shape" >> $O/r2n_probe.log
  timeout 300 python scripts/perf_probe2.py $shape rq_min_tiles=$NEVER rq_min_tiles=64 >> $O/r2n_probe.log 2>&1
done
cat $O/r2n_probe.log
M=sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,l1tex__m_xbar2l1tex_read_bytes.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,smsp__cycles_active.avg,sm__cycles_elapsed.max
for rq in $NEVER 64; do
TVC_RQ_MIN_TILES=$rq timeout 300 ncu --metrics $M --clock-control none -k regex:gemm_topk_pair --launch-skip 1 --launch-count 1 --csv --log-file $O/r2n_ncu_rq$rq.csv python scripts/perf_probe2.py 81920x1000000x448 default > /dev/null 2>&1
grep -v "^==" $O/r2n_ncu_rq$rq.csv | awk -F'","' 'NR>1{print $5, $13, $15}' | cut -c1-200
done
