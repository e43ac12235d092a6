#!/bin/bash
mkdir -p gpurun_out
for v in 1 2 4 8; do
TNSB_BUCKET_PASSES=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_p$v.json 2> gpurun_out/bench_p$v.err; tail -c 300 gpurun_out/bench_p$v.err
done
