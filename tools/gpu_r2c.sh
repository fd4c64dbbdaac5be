#!/bin/bash
# level-set exactness at large sizes (every call bounded: a wavefront bug shows as a hang)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "== ls_diag 4096 step 0"; timeout 150 python tools/ls_diag.py 4096 0 2>&1 | tail -12; echo "rc=$?"
echo "== ls_diag 1024 after 1 step dt 0.005"; timeout 100 python tools/ls_diag.py 1024 1 0.005 2>&1 | tail -8
echo "== ls_diag 2048 after 2 steps"; timeout 150 python tools/ls_diag.py 2048 2 2>&1 | tail -8
if [[ "$1" == "noskip" ]]; then
  echo "== ls_diag 4096 step 0 NOSKIP"; FSIM_LS_NOSKIP=1 timeout 150 python tools/ls_diag.py 4096 0 2>&1 | tail -12
fi
