#!/bin/bash
# A/B of the two query kernels on one B200: parity first, then bench lines for both (gpurun_out/).
set -o pipefail
mkdir -p gpurun_out
echo "== quick parity (round kernel)" 
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "golden or one_set or two_dynamic or c1_100k or long_lists or overflow or query_limit or empty or 64bit" 2>&1 | tail -15 | tee gpurun_out/quick.log
if ! grep -q "passed" gpurun_out/quick.log || grep -q "failed" gpurun_out/quick.log; then echo "QUICK PARITY FAILED"; fi
echo "== bench round kernel"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_rounds.json 2> gpurun_out/bench_rounds.err; tail -c 3000 gpurun_out/bench_rounds.json
echo "== bench cell kernel"
TNSB_QUERY_KERNEL=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cells.json 2> gpurun_out/bench_cells.err; tail -c 1500 gpurun_out/bench_cells.json
echo "== full gpu suite (round kernel)"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/full.log
