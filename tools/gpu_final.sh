#!/bin/bash
# final validation of a candidate state: full GPU parity suite + default bench (configs 2, 3, 4, 80M, micro-benchmark in one line) (gpurun_out/)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/full.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 400 gpurun_out/bench_default.err
