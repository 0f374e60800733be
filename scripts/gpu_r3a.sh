#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_consistency.py tests/test_gpu_api.py -x -q > $O/r3a_tests.log 2>&1; tail -2 $O/r3a_tests.log
timeout 300 python scripts/bench_bc.py 2>&1 | grep -E "emb|sims" > $O/r3a_bench_bc.log; cat $O/r3a_bench_bc.log
python scripts/trace_emb.py 2>&1 | sed -n 1,3p; python scripts/trace_emb.py 2>&1 | tail -8
