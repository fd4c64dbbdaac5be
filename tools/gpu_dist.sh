#!/bin/bash
# multi-GPU session: oracle parity of the y-slab projection, then the bench line at N GPUs
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
if [[ "$2" != "benchonly" ]]; then
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py ${DIST_SIZE:-256} 3 > gpurun_out/dist_check_$N.log 2>&1; echo "dist_check rc=$?"
  grep -E "slab parity|rank 0 step|n=|OK|Error|error|assert" gpurun_out/dist_check_$N.log | tail -20
fi
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps ${STEPS:-10} --warmup ${WARMUP:-5} ${BENCH_ARGS} > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_${N}gpu.err
python tools/bench_summary.py gpurun_out/bench_${N}gpu.json
