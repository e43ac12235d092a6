#!/bin/bash
# round-2 evidence (run under gpurun, ONE GPU): launch list of a short bench.py run + full ncu capture of the dominant kernel
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1
# 4th launch of the query kernel: the steady-state variant (the first run uses the tall hit columns)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:brick_query -s 3 -c 1 -f -o gpurun_out/r2_brick_query \
    python tools/run_workload.py --steps 5 > gpurun_out/r2_ncu_brick_query.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bucket_scatter -s 9 -c 1 -f -o gpurun_out/r2_scatter \
    python tools/run_workload.py --steps 5 > gpurun_out/r2_ncu_scatter.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:brick_query -s 3 -c 1 -f -o gpurun_out/r2_brick_query_dambreak \
    python tools/run_workload.py --steps 5 --workload dambreak > gpurun_out/r2_ncu_brick_query_dambreak.log 2>&1
tail -1 gpurun_out/r2_ncu_brick_query.log | cut -c1-400
