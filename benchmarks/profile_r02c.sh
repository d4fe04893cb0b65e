#!/bin/bash
# third GPU pass of round 2: full GPU parity suite, the driver's bench command, and a source-level ncu capture of the
# headline kernel at HEAD
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest_gpu.log
tail -3 gpurun_out/r02c_pytest_gpu.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02c_bench_line.json 2> gpurun_out/r02c_bench.err; echo "bench rc=$?"
NCU="ncu --clock-control none"
timeout 600 $NCU --set full --import-source on -k regex:opd_highway_multi -s 1 -c 1 -f -o gpurun_out/r02c_opd_multi \
    python bench.py --steps 1 --warmup 3 --trees 4736 --headline-only --no-cpu-baseline > gpurun_out/r02c_prof_multi.log 2>&1
ls -la gpurun_out/
