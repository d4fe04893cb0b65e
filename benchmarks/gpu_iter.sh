#!/bin/bash
# one GPU iteration while tuning hw::step: env-step parity first (fail fast), then the whole GPU suite, then the
# headline throughput of the default library and of any variant libraries under build/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-iter}
timeout 600 python -m pytest tests/test_gpu_engines.py -x -q -m gpu -k "highway or constant_divisor" > gpurun_out/${TAG}_pytest_step.log 2>&1
echo "step tests rc=$?"; tail -2 gpurun_out/${TAG}_pytest_step.log
if [ -z "$SKIP_FULL" ]; then timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1; fi
echo "gpu suite rc=$?"; tail -2 gpurun_out/${TAG}_pytest_gpu.log 2>/dev/null
timeout 300 python bench.py --steps 5 --warmup 3 --headline-only --no-cpu-baseline > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
python -c "import json;d=json.loads(open('gpurun_out/${TAG}_bench_default.json').read().strip().split('\n')[-1]);print('default', d['value'], d['ms_per_step'])"
for lib in build/libb2planner_*.so; do
  [ -f "$lib" ] || continue
  n=$(basename $lib .so)
  B2PLANNER_LIB=$PWD/$lib timeout 300 python bench.py --steps 5 --warmup 3 --headline-only --no-cpu-baseline > gpurun_out/${TAG}_bench_$n.json 2> gpurun_out/${TAG}_bench_$n.err
  python -c "import json;d=json.loads(open('gpurun_out/${TAG}_bench_$n.json').read().strip().split('\n')[-1]);print('$n', d['value'], d['ms_per_step'])"
done
if [ -n "$WITH_NCU" ]; then
  timeout 600 ncu --clock-control none --set full --import-source on -k regex:opd_highway_multi -s 1 -c 1 -f -o gpurun_out/${TAG}_opd_multi \
    python bench.py --steps 1 --warmup 3 --trees 4736 --headline-only --no-cpu-baseline > gpurun_out/${TAG}_prof_multi.log 2>&1
  ls -la gpurun_out/${TAG}_opd_multi.ncu-rep
fi
