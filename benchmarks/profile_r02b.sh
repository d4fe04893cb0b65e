#!/bin/bash
# second ncu pass of round 2: the headline kernel after the hw::step changes, and the speculative strict kernel
set -x
cd "$(dirname "$0")/.."
NCU="ncu --clock-control none"
$NCU --set full --import-source on -k regex:opd_highway_multi -s 1 -c 1 -f -o gpurun_out/r02b_opd_multi \
    python bench.py --steps 1 --warmup 3 --trees 4736 --headline-only --no-cpu-baseline > gpurun_out/r02b_prof_multi.log 2>&1
$NCU --set full --import-source on -k regex:opd_spec -s 1 -c 1 -f -o gpurun_out/r02b_opd_spec \
    python benchmarks/bench_spec.py --widths 256 --strict 0 --seeds 1 --reps 1 > gpurun_out/r02b_prof_spec.log 2>&1
python benchmarks/bench_spec.py --budgets 3125,10000,15625 --gammas 0.8,0.95 > gpurun_out/r02b_spec.json 2> gpurun_out/r02b_spec.err
ls -la gpurun_out/*.ncu-rep
