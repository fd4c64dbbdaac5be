#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "capture/" \
    -k regex:"axpyKernel|updateVelocityKernel|g2pKernel|advectKernel|lsBinKernel|lsRelabelStatsKernel|sdPackKernel" \
    -c 24 -o /tmp/r2_full2 -f python tools/ncu_target.py 4096 3 3 > gpurun_out/ncu_full2.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r2_full2.ncu-rep --page raw --csv > gpurun_out/r2_full2_raw.csv 2>/dev/null
gzip -f gpurun_out/r2_full2_raw.csv
tail -2 gpurun_out/ncu_full2.log
