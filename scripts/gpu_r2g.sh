#!/bin/bash
# round-2 GPU session g: batched kernel (c), TMA-staged kernel (b) statistics, paced kernel (a) (A/B + DRAM traffic)
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2g_gpu_suite.log 2>&1; echo "pytest rc=$?" >> $O/r2g_gpu_suite.log
timeout 300 python scripts/bench_bc.py > $O/r2g_bench_bc.log 2>&1
for agg in 0 1; do
  echo "== TVC_KOCC_AGG=$agg" >> $O/r2g_bench_bc.log
  TVC_KOCC_AGG=$agg timeout 200 python scripts/bench_bc.py quick 2>&1 | grep kocc >> $O/r2g_bench_bc.log
done
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
for pace in "16 3" "0 3" "8 2" "32 4" "16 3" "0 3"; do
  set -- $pace
  echo "== TVC_PACE_EVERY=$1 TVC_PACE_AHEAD=$2" >> $O/r2g_pace.log
  TVC_PACE_EVERY=$1 TVC_PACE_AHEAD=$2 timeout 300 $B 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print(json.dumps({'value':j['value'],'ms_per_step':j['ms_per_step'],'kernel_ms':j['roofline']['kernel_ms_per_step'],'frac':j['roofline']['frac'],'digest':j['parity']['digest'],'clk':j['clocks']['sm_mhz']}))
" >> $O/r2g_pace.log
done
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum
for pace in 16 0; do
  TVC_PACE_EVERY=$pace timeout 600 ncu --metrics $M --clock-control none -k regex:gemm_topk_pair --launch-skip 2 --launch-count 2 --csv \
     --log-file $O/r2g_traffic_pace$pace.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras --no-parity > $O/r2g_traffic_pace$pace.log 2>&1
done
tail -3 $O/r2g_gpu_suite.log; cat $O/r2g_bench_bc.log $O/r2g_pace.log; grep -v "^==" $O/r2g_traffic_pace16.csv | tail -12; grep -v "^==" $O/r2g_traffic_pace0.csv | tail -12
