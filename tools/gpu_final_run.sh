set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; tail -c 1500 gpurun_out/bench_r02.json
ncu --set full --clock-control none --import-source on -k regex:noise_grid2 -s 2 -c 1 -f -o gpurun_out/prof_noise2_r02 python bench.py --steps 1 --warmup 3 --kernel-only > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r02_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
ls -la gpurun_out | tail -5
