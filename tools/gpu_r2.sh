#!/bin/bash
# one GPU session of round 2: parity tests, the bench line, then the ncu evidence (launch list + full captures)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
WHAT=${1:-all}
if [[ $WHAT == all || $WHAT == tests ]]; then
  timeout 900 python -m pytest tests -m gpu -q -rA --tb=short --timeout 420 ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
  grep -E "^(PASSED|FAILED|ERROR|SKIPPED)|passed|failed|Error|assert|cap test|config [12]|4096\^2" gpurun_out/pytest_gpu.log | tail -80
fi
if [[ $WHAT == all || $WHAT == bench ]]; then
  timeout 300 python bench.py --steps 10 --warmup 5 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"
  python tools/bench_summary.py gpurun_out/bench_1gpu.json
fi
if [[ $WHAT == ab ]]; then
  FSIM_FUSED_AXPY=1 timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu --no-e2e > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err; echo "bench fused rc=$?"
  python tools/bench_summary.py gpurun_out/bench_fused.json
fi
if [[ $WHAT == all || $WHAT == ncu ]]; then
  # launch list of one bench step (cold-cache, serialised: shares only)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches.csv \
      python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
  # full captures of one instance of each kernel class (third step of a run capped at 3 PCG iterations)
  timeout 1200 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "capture/" \
      -k regex:"${NCU_KERNELS:-applyASdKernel|solveKernelR|assembleKernel|p2gGatherKernel|g2pKernel|advectKernel|layerFillKernel|updateVelocityKernel|sweepKernel}" \
      -c ${NCU_COUNT:-60} -o /tmp/r2_full -f python tools/ncu_target.py 4096 3 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
  # the report stays on the box (hundreds of MB); what travels back are the per-kernel metric tables
  ncu -i /tmp/r2_full.ncu-rep --page raw --csv > gpurun_out/r2_full_raw.csv 2>/dev/null
  ncu -i /tmp/r2_full.ncu-rep --page details --csv > gpurun_out/r2_full_details.csv 2>/dev/null
  gzip -f gpurun_out/r2_full_raw.csv gpurun_out/r2_full_details.csv gpurun_out/r2_launches.csv
  tail -3 gpurun_out/ncu_full.log
  ls -la gpurun_out/
fi
