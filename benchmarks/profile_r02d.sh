#!/bin/bash
# final GPU pass of round 2: whole GPU parity suite, smoke(), the driver's bench command, and the ncu launch list of
# the bench command
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest_gpu.log
tail -3 gpurun_out/r02d_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02d_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02d_bench_line.json 2> gpurun_out/r02d_bench.err; echo "bench rc=$?"
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02d_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02d_launches_bench.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/ | grep r02d
