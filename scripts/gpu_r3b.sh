#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_search.py -x -q -k "resident_query_kernels_match_pair_kernel and (4500 or 700 or 513)" > $O/r3b_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/r3b_memcheck.log
tail -6 $O/r3b_memcheck.log
for mb in 4 6 8; do echo "== TVC_SIMS_MIN_BLOCKS=$mb"; TVC_SIMS_MIN_BLOCKS=$mb timeout 200 python scripts/bench_bc.py quick 2>&1 | grep -E "sims"; done > $O/r3b_sims.log 2>&1
cat $O/r3b_sims.log
