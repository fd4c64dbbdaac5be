#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
for sg in 2 3; do
  echo "=== FSIM_SD_SIGMA=$sg: parity subset"
  FSIM_SD_SIGMA=$sg timeout 600 python -m pytest tests/test_gpu_parity.py -q --timeout 300 -k "stagewise or trajectory or ragged or random_divergence or deterministic or full_size_projection or config2" 2>&1 | tail -4
  echo "=== FSIM_SD_SIGMA=$sg: bench"
  FSIM_SD_SIGMA=$sg timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_sigma$sg.json 2>/dev/null
  python tools/bench_summary.py gpurun_out/bench_sigma$sg.json | grep -E "value|mic0|applyA|axpy|stages"
  FSIM_SD_SIGMA=$sg timeout 300 python bench.py --workload projection --warmup 1 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); print('  ', d['metric'], 'iters', d['iterations'], 'iter/s %.0f' % d['value'], 'ms %.1f' % d['ms_projection'])"
done
