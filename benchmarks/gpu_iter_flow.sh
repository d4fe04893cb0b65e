#!/bin/bash
# GPU iteration for the dataflow OPD kernel: parity (incl. C2 full size), then throughput of both batch kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-flow}
timeout 900 python -m pytest tests/test_gpu_engines.py -x -q -m gpu -k "highway or constant_divisor" > gpurun_out/${TAG}_pytest_step.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_pytest_step.log
for k in 0 2; do
  timeout 300 python bench.py --steps 5 --warmup 3 --headline-only --no-cpu-baseline --kernel $k > gpurun_out/${TAG}_bench_k$k.json 2> gpurun_out/${TAG}_bench_k$k.err
  python -c "import json;d=json.loads(open('gpurun_out/${TAG}_bench_k$k.json').read().strip().split('\n')[-1]);print('kernel $k', d['value'], d['ms_per_step'])"
done
for lib in build/libb2planner_*.so; do
  [ -f "$lib" ] || continue
  n=$(basename $lib .so)
  for k in 0 2; do
    B2PLANNER_LIB=$PWD/$lib timeout 300 python bench.py --steps 5 --warmup 3 --headline-only --no-cpu-baseline --kernel $k > gpurun_out/${TAG}_bench_${n}_k$k.json 2> gpurun_out/${TAG}_bench_${n}_k$k.err
    python -c "import json;d=json.loads(open('gpurun_out/${TAG}_bench_${n}_k$k.json').read().strip().split('\n')[-1]);print('$n kernel $k', d['value'], d['ms_per_step'])"
  done
done
