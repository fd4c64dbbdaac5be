ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:solveKernelR -s 40 -c 2 -f -o gpurun_out/solve_r1 python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/solve_r1.ncu-rep --page raw --csv > gpurun_out/solve_r1_raw.csv 2>/dev/null
gzip -kf gpurun_out/launches_r1.csv
(time python bench.py) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -4 gpurun_out/bench_default.err; head -c 400 gpurun_out/bench_default.json
