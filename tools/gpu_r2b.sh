#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "== ls_diag 4096 step 0"; timeout 600 python tools/ls_diag.py 4096 0 2>&1 | tail -25
echo "== ls_diag 4096 step 0 LEGACY"; FSIM_LS_LEGACY=1 timeout 600 python tools/ls_diag.py 4096 0 2>&1 | tail -25
echo "== ls_diag 1024 after 1 step dt 0.005"; timeout 600 python tools/ls_diag.py 1024 1 0.005 2>&1 | tail -25
echo "== ls_diag 1024 after 1 step dt 0.005 LEGACY"; FSIM_LS_LEGACY=1 timeout 600 python tools/ls_diag.py 1024 1 0.005 2>&1 | tail -25
echo "== mirror roundtrip x6"
for k in 1 2 3 4 5 6; do timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "mirror_roundtrip and semilagrangian" 2>&1 | tail -2; done
for d in 1 2 3; do
  echo "== fused bench FSIM_DBG_PRE=$d"
  FSIM_DBG_PRE=$d timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_dbg$d.json 2>/dev/null
  python tools/bench_summary.py gpurun_out/bench_dbg$d.json | grep -E "value|mic0|applyA"
done
