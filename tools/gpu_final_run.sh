#!/bin/bash
# The GPU-side sequence behind profiles/: full GPU test suite, smoke, the default bench line, the launch list, and (with "ncu" as the first argument) the
# --set full capture of the headline kernel.   gpurun --timeout 1800 -- 'bash tools/gpu_final_run.sh [ncu]'
set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; tail -c 1800 gpurun_out/bench_r02.json; tail -3 gpurun_out/bench_r02.err
if [ "$1" = "ncu" ]; then
  ncu --set full --clock-control none --import-source on -k regex:noise_grid2 -s 2 -c 1 -f -o gpurun_out/prof_noise2_r02 python bench.py --steps 1 --warmup 3 --kernel-only > /dev/null 2>&1
fi
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r02_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
ls -la gpurun_out | tail -5
