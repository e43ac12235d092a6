#!/bin/bash
mkdir -p gpurun_out
for wl in dambreak; do
  for k in 1 0; do
    TNSB_QUERY_KERNEL=$k timeout 900 python bench.py --steps 5 --warmup 3 --workload $wl --no-cpu-baseline > gpurun_out/bench_${wl}_k$k.json 2> gpurun_out/bench_${wl}_k$k.err
    tail -c 400 gpurun_out/bench_${wl}_k$k.err
  done
done
