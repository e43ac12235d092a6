#!/bin/bash
# parity first, then bench lines for the A/B switches (gpurun_out/).
set -o pipefail
mkdir -p gpurun_out
echo "== full gpu suite"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/full.log
echo "== bench default (bucket build, cell kernel)"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.err
echo "== bench radix build"
TNSB_BUILD=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_radix.json 2> gpurun_out/bench_radix.err; tail -c 600 gpurun_out/bench_radix.err
echo "== bench round kernel"
TNSB_QUERY_KERNEL=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_rounds.json 2> gpurun_out/bench_rounds.err; tail -c 600 gpurun_out/bench_rounds.err
echo "== dambreak default"
timeout 600 python bench.py --steps 5 --warmup 3 --workload dambreak --no-cpu-baseline > gpurun_out/bench_dambreak.json 2> gpurun_out/bench_dambreak.err; tail -c 600 gpurun_out/bench_dambreak.err
