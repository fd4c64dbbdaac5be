python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1b_tests.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r1b_bench.json 2> gpurun_out/r1b_bench.err
tail -3 gpurun_out/r1b_tests.log; cat gpurun_out/r1b_bench.json; tail -5 gpurun_out/r1b_bench.err
