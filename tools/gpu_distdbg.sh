#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
N=${1:-2}
for d in 0 16 32 112; do
  echo "=== FSIM_DBG_PRE=$d"
  FSIM_DBG_PRE=$d timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$((d % 7)) bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --no-e2e --no-parity > gpurun_out/bench_dbg_$d.json 2>/dev/null
  python tools/bench_summary.py gpurun_out/bench_dbg_$d.json 2>/dev/null | grep -E "value|mic0_f|mic0_b|applyA|axpy" 
done
