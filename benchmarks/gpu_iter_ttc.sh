#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_ttc_vi.py -q -m gpu > gpurun_out/ttc_pytest.log 2>&1; echo "ttc tests rc=$?"; tail -15 gpurun_out/ttc_pytest.log
timeout 300 python -m pytest tests/test_gpu_agents.py tests/test_gpu_cabi.py -x -q -m gpu > gpurun_out/ttc_pytest_agents.log 2>&1; echo "agents rc=$?"; tail -2 gpurun_out/ttc_pytest_agents.log
timeout 120 python benchmarks/bench_ttc_vi.py > gpurun_out/ttc_bench.json 2> gpurun_out/ttc_bench.err; echo "bench rc=$?"; cat gpurun_out/ttc_bench.json
