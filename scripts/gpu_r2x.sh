#!/bin/bash
# round-2 GPU session x: ncu launch list + DRAM traffic of the default tree (resident-query pair kernel)
O=gpurun_out; mkdir -p $O
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/r2x_bench_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2x_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/r2x_ncu.log 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,l1tex__m_xbar2l1tex_read_bytes.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:gemm_topk_pair --launch-skip 2 --launch-count 2 --csv --log-file $O/r2x_traffic.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras --no-parity > $O/r2x_traffic.log 2>&1
grep -v "^==" $O/r2x_traffic.csv | awk -F'","' 'NR>1{print substr($5,1,60), $13, $15}'
tail -c 600 $O/r2x_bench_plain.log
