for N in 2 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 3 --warmup 2 --no-cpu > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "N=$N rc $?"
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_n$N.json') if l.startswith('{')][-1])
    print($N, d['value'], d['ms_per_step'], d['config']['stage_ms_last_step'], d['config']['pcg_iters_last_step'], d['config'].get('pcg_residual_last_step'), d.get('e2e'))
    for k,v in d['config']['kernels'].items(): print('   ',k, round(v['avg_ms'],4))
except Exception as e:
    print('fail', e); print(open('gpurun_out/bench_n$N.err').read()[-2000:])
PY
done
