#!/bin/bash
# throughput of trees-per-CTA variants of the batch kernel (bench only)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-var}
run() { # name lib trees extra
  B2PLANNER_LIB=$2 timeout 300 python bench.py --steps 4 --warmup 3 --headline-only --no-cpu-baseline --trees $3 $4 > gpurun_out/${TAG}_bench_$1.json 2> gpurun_out/${TAG}_bench_$1.err
  python -c "import json;d=json.loads(open('gpurun_out/${TAG}_bench_$1.json').read().strip().split('\n')[-1]);print('$1', d['value'], d['ms_per_step'], d['run_config']['trees_per_gpu'])"
}
D=$PWD/rl_agents_b200/csrc/libb2planner.so
run t8 $D 18944 ""
run t8_smemkeys $D 18944 "--keys-in-smem 1"
run t16 $PWD/build/libb2planner_t16.so 18944 ""
run t12 $PWD/build/libb2planner_t12.so 14208 ""
run t10 $PWD/build/libb2planner_t10.so 17760 ""
run t6 $PWD/build/libb2planner_t6.so 17760 ""
