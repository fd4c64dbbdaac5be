timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/b.json 2> gpurun_out/b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/b.json'))
print(d['value'], d['ms_per_step'], [round(x,1) for x in d['config']['stage_ms_last_step']], d['config']['pcg_iters_last_step'], d['config']['pcg_residual_last_step'])
for k,v in d['config']['kernels'].items(): print('  ',k, v['launches'], round(v['avg_ms'],4))
PY
tail -3 gpurun_out/b.err
FSIM_SD_RPL=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('RPL=1:', d['value'], d['ms_per_step'])"
