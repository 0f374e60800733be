#!/bin/bash
# round-2 GPU session f: new kernel (c), variants of (b)/(c), ncu source-level captures, cluster occupancy probe
mkdir -p gpurun_out
O=gpurun_out
scripts/build/cluster_probe > $O/r2f_cluster_probe.json 2>&1
timeout 300 python -m pytest tests/test_gpu_hubness.py -x -q > $O/r2f_hubness.log 2>&1; echo "rc=$?" >> $O/r2f_hubness.log
for agg in default 0 1; do
  if [ $agg = default ]; then unset TVC_KOCC_AGG; else export TVC_KOCC_AGG=$agg; fi
  echo "== TVC_KOCC_AGG=$agg" >> $O/r2f_bench_bc.log
  timeout 200 python scripts/bench_bc.py quick 2>&1 | grep kocc >> $O/r2f_bench_bc.log
done
unset TVC_KOCC_AGG
for mb in 4 6 8; do
  echo "== TVC_SIMS_MIN_BLOCKS=$mb" >> $O/r2f_bench_bc.log
  TVC_SIMS_MIN_BLOCKS=$mb timeout 200 python scripts/bench_bc.py quick 2>&1 | grep -E "sims|emb " >> $O/r2f_bench_bc.log
done
echo "== full" >> $O/r2f_bench_bc.log
timeout 300 python scripts/bench_bc.py >> $O/r2f_bench_bc.log 2>&1
for kn in consistency_sims_kernel k_occurrence_kernel consistency_emb_pipe_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$kn --launch-skip 3 --launch-count 1 \
    -o $O/r2f_$kn -f python scripts/bench_bc.py quick > $O/r2f_ncu_$kn.log 2>&1
done
cat $O/r2f_cluster_probe.json; tail -3 $O/r2f_hubness.log; cat $O/r2f_bench_bc.log
