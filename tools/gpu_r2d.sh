#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "== ls_diag 4096 step 0"; timeout 150 python tools/ls_diag.py 4096 0 2>&1 | grep -E "sweeps|differ"
echo "== ls_diag 4096 step 0 NOSKIP"; FSIM_LS_NOSKIP=1 timeout 150 python tools/ls_diag.py 4096 0 2>&1 | grep -E "sweeps|differ"
echo "== ls_diag 2048 after 3 steps"; timeout 200 python tools/ls_diag.py 2048 3 2>&1 | grep -E "sweeps|differ"
echo "== ls_diag 2048 after 3 steps NOSKIP"; FSIM_LS_NOSKIP=1 timeout 200 python tools/ls_diag.py 2048 3 2>&1 | grep -E "sweeps|differ"
echo "== config 1 (128^2 semi-Lagrangian)"
timeout 300 python bench.py --workload sl --size 128 > gpurun_out/bench_sl_128.json 2> gpurun_out/bench_sl_128.err
python - <<'PY'
import json
for ln in open('gpurun_out/bench_sl_128.json'):
    if ln.startswith('{'):
        d = json.loads(ln); print('  ', d['config']['workload'][:110], '| ms/step %.2f' % d['ms_per_step'], '| stages', [round(x, 2) for x in d['config']['stage_ms_last_step']], '| cpu', (d.get('cpu_baseline') or {}).get('ms_per_step'))
PY
echo "== config 3 (projection stress)"
timeout 300 python bench.py --workload projection --warmup 1 > gpurun_out/bench_projection.json 2> gpurun_out/bench_projection.err
python - <<'PY'
import json
for ln in open('gpurun_out/bench_projection.json'):
    if ln.startswith('{'):
        d = json.loads(ln); print('  ', d['metric'], 'iters', d['iterations'], 'iter/s %.0f' % d['value'], 'ms %.1f' % d['ms_projection'], 'frac %.3f' % d['hbm_frac_algorithmic'])
PY
