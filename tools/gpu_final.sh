#!/bin/bash
# final GPU session of round 2: the whole -m gpu suite, the bench lines of configs 1, 2, 3 and the headline, the ncu launch list
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -rA --tb=short --timeout 420 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR" gpurun_out/pytest_gpu.log | tail -8
timeout 300 python bench.py --steps 10 --warmup 5 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"
python tools/bench_summary.py gpurun_out/bench_1gpu.json 2>/dev/null | head -3
timeout 300 python bench.py --size 1024 --steps 20 --warmup 5 > gpurun_out/bench_1024_config2.json 2> gpurun_out/bench_1024.err; echo "config2 rc=$?"
python tools/bench_summary.py gpurun_out/bench_1024_config2.json 2>/dev/null | head -2
timeout 300 python bench.py --workload sl --size 128 > gpurun_out/bench_sl_128.json 2> gpurun_out/bench_sl_128.err; echo "config1 rc=$?"
timeout 300 python bench.py --workload projection --warmup 1 > gpurun_out/bench_projection.json 2> gpurun_out/bench_projection.err; echo "config3 rc=$?"
python - <<'PY'
import json
for fn in ('gpurun_out/bench_sl_128.json', 'gpurun_out/bench_projection.json'):
    for ln in open(fn):
        if ln.startswith('{'):
            d = json.loads(ln); print('  ', d['metric'][:60], '| value %.4g' % d['value'], '| ms/step', d.get('ms_per_step'), d.get('ms_projection'))
PY
if [[ "$1" == ncu ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2b_launches.csv \
      python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
  gzip -f gpurun_out/r2b_launches.csv
fi
