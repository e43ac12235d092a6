#!/bin/bash
# round-end evidence: launch list of one bench.py run + full ncu capture of the dominant kernel (gpurun_out/)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:query_kernel -s 1 -c 1 -f -o gpurun_out/query_kernel_final python tools/run_workload.py --steps 3 > gpurun_out/ncu_query_final.log 2>&1
tail -2 gpurun_out/ncu_query_final.log
