timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "N=2 rc $?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_n2.json') if l.startswith('{')][-1])
print(2, d['value'], d['ms_per_step'], [round(x,1) for x in d['config']['stage_ms_last_step']], d['config']['pcg_iters_last_step'], d['config'].get('pcg_residual_last_step'), d.get('e2e',{}).get('value'))
for k,v in d['config']['kernels'].items(): print('   ',k, v['launches'], round(v['avg_ms'],4))
PY
tail -3 gpurun_out/bench_n2.err
timeout 800 python bench.py --workload projection --warmup 1 > gpurun_out/proj.json; python -c "
import json
for l in open('gpurun_out/proj.json'):
    d=json.loads(l); print(d['metric'], round(d['value'],1), d['iterations'], round(d['ms_projection'],1), round(d['hbm_frac_algorithmic'],3))"
