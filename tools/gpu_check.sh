#!/bin/bash
# What a round-end check runs on a GPU box (via gpurun): parity tests, smoke, the default bench.
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
(time python bench.py) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -4 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline'], d['cpu_baseline']['value'], d['clocks'], d['gpu_launches'])
print(d['config']['step_dense_model_frac'], d['config']['step_hbm_frac_touched'], d['config']['pcg_iter_per_s'])
for k,v in d['config']['kernels'].items(): print('   ',k, v['launches'], round(v['avg_ms'],4), round(v.get('gbs',0)))
PY
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 | head -c 600
