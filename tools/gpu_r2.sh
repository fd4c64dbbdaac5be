#!/bin/bash
# one GPU session of round 2: parity tests, the bench line, then the ncu evidence (launch list + full captures)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
WHAT=${1:-all}
if [[ $WHAT == all || $WHAT == tests ]]; then
  timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
  tail -5 gpurun_out/pytest_gpu.log
fi
if [[ $WHAT == all || $WHAT == bench ]]; then
  timeout 600 python bench.py --steps 10 --warmup 5 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"
  cat gpurun_out/bench_1gpu.json | head -c 6000
fi
if [[ $WHAT == all || $WHAT == ncu ]]; then
  # launch list of one bench step (cold-cache, serialised: shares only)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches.csv \
      python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
  # full captures of one instance of each kernel class (third step of a run capped at 3 PCG iterations)
  timeout 1200 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "capture/" \
      -k regex:'applyASdKernel|solveKernelR|assembleKernel|p2gGatherKernel|g2pKernel|advectKernel|layerFillKernel|updateVelocityKernel|pcgFinishKernel|sweepKernel|lsBinKernel|deriveKernel|sdPackKernel' \
      -c ${NCU_COUNT:-120} -o gpurun_out/r2_full -f python tools/ncu_target.py 4096 3 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
  tail -3 gpurun_out/ncu_full.log
  ls -la gpurun_out/
fi
