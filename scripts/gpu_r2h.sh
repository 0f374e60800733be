#!/bin/bash
# round-2 GPU session h: pacing window sweep for kernel (a), cheap parked waits in kernel (b) emb
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_consistency.py tests/test_gpu_api.py -x -q > $O/r2h_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2h_tests.log
timeout 300 python scripts/bench_bc.py > $O/r2h_bench_bc.log 2>&1
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
for pace in "8 2" "0 1" "4 2" "8 1" "4 4" "16 2" "8 3" "2 4" "8 2" "0 1"; do
  set -- $pace
  echo "== TVC_PACE_EVERY=$1 TVC_PACE_AHEAD=$2" >> $O/r2h_pace.log
  TVC_PACE_EVERY=$1 TVC_PACE_AHEAD=$2 timeout 300 $B 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print(json.dumps({'value':j['value'],'ms_per_step':j['ms_per_step'],'kernel_ms':j['roofline']['kernel_ms_per_step'],'frac':j['roofline']['frac'],'digest':j['parity']['digest'],'clk':j['clocks']['sm_mhz']}))
" >> $O/r2h_pace.log
done
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum
for pace in "8 2" "4 2" "8 1"; do
  set -- $pace
  TVC_PACE_EVERY=$1 TVC_PACE_AHEAD=$2 timeout 600 ncu --metrics $M --clock-control none -k regex:gemm_topk_pair --launch-skip 2 --launch-count 1 --csv \
     --log-file $O/r2h_traffic_pace$1_$2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras --no-parity > $O/r2h_traffic.log 2>&1
  echo "== pace $1 $2" >> $O/r2h_traffic_summary.log
  grep -v "^==" $O/r2h_traffic_pace$1_$2.csv | awk -F'","' 'NR>1{print $13, $15}' >> $O/r2h_traffic_summary.log
done
tail -3 $O/r2h_tests.log; grep -E "emb|sims" $O/r2h_bench_bc.log; cat $O/r2h_pace.log $O/r2h_traffic_summary.log
