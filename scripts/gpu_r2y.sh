#!/bin/bash
# multi-GPU session: multi-GPU parity test + bench at the box's GPU count
O=gpurun_out; mkdir -p $O
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > $O/r2y_multi_test_n$N.log 2>&1; tail -3 $O/r2y_multi_test_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2y_bench_n$N.log 2>&1
python - <<PY
import json
for l in open('gpurun_out/r2y_bench_n$N.log'):
    if l.startswith('{'):
        j=json.loads(l)
        print(json.dumps({k:j[k] for k in ('value','ms_per_step','n_gpus')}), 'e2e', j['e2e']['value'], 'pageable', j['e2e']['pageable']['value'], 'digest', j['parity']['digest'], 'frac', j['roofline']['frac'], 'share', j['roofline']['kernel_share_of_step'])
PY
