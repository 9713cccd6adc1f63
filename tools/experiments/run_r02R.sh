#!/bin/bash
# r02R batch (one gpurun call): final-build evidence — GPU test suite, smoke, bench line (20 steps as the driver runs it),
# reference arm, and the grouped-lane reduction microbenchmark.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02R_pytest_gpu.txt 2>&1; tail -3 gpurun_out/r02R_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02R_smoke.txt 2>&1; tail -2 gpurun_out/r02R_smoke.txt
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02R_bench_reference_c3.json 2> gpurun_out/r02R_bench_reference.err; cut -c1-600 gpurun_out/r02R_bench_reference_c3.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02R_bench_c3_20steps.json 2> gpurun_out/r02R_bench.err; cut -c1-700 gpurun_out/r02R_bench_c3_20steps.json; tail -3 gpurun_out/r02R_bench.err
timeout 300 python tools/experiments/red_coalescing.py --out gpurun_out/r02Q_exp_red_coalescing.json 2>&1 | cut -c1-300
