#!/bin/bash
# ncu evidence of round 2 (run on the GPU box through gpurun; outputs under gpurun_out/)
set -x
cd "$(dirname "$0")/.."
NCU="ncu --clock-control none"
# 1. every launch of a short bench run with its device time
$NCU --metrics gpu__time_duration.sum -c 80 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --trees 4736 --headline-only --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1
# 2. the headline kernel (32 trees per SM: the same kernel and occupancy, a shorter launch for the ~40 replays)
$NCU --set full --import-source on -k regex:opd_highway_multi -s 1 -c 1 -f -o gpurun_out/r02_opd_multi \
    python bench.py --steps 1 --warmup 3 --trees 4736 --headline-only --no-cpu-baseline > gpurun_out/r02_prof_multi.log 2>&1
# 3. value iteration, C4 shape, the shipped row kernel
$NCU --set full --import-source on -k regex:vi_sweep_row -s 10 -c 1 -f -o gpurun_out/r02_vi_row \
    python benchmarks/bench_vi.py --sweeps 20 --cpu-sweeps 0 > gpurun_out/r02_prof_vi.log 2>&1
# 4. one C2 decision in waves of 64 leaves
$NCU --set full --import-source on -k regex:opd_wave -s 1 -c 1 -f -o gpurun_out/r02_opd_wave \
    python benchmarks/bench_wave.py --widths 64 --strict 0 --seeds 1 --reps 1 > gpurun_out/r02_prof_wave.log 2>&1
# 5. one C3 decision in waves of 512 episodes
$NCU --set full --import-source on -k regex:mcts_wave -s 1 -c 1 -f -o gpurun_out/r02_mcts_wave \
    python benchmarks/bench_mcts_wave.py > gpurun_out/r02_prof_mcts.log 2>&1
ls -la gpurun_out/*.ncu-rep
