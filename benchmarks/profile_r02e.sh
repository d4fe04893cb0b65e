#!/bin/bash
# closing GPU pass of round 2 at HEAD: whole GPU parity suite, smoke(), a short bench line with every other_paths key
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/r02e_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02e_pytest_gpu.log
tail -3 gpurun_out/r02e_pytest_gpu.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02e_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02e_smoke.log
timeout 110 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_bench_line.json 2> gpurun_out/r02e_bench.err; echo "bench rc=$?"
python -c "import json;d=json.loads(open('gpurun_out/r02e_bench_line.json').read().strip().split('\n')[-1]);print(d['value'], json.dumps(d['other_paths'].get('vi_highway_ttc'))[:400])"
