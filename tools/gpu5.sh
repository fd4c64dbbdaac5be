timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 256 3 > gpurun_out/dist256.log 2>&1; echo "rc $?" >> gpurun_out/dist256.log
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/dist256.log | tail -12
for N in 2 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 3 --warmup 2 --no-cpu > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "N=$N rc $?"
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_n$N.json') if l.startswith('{')][-1])
    print($N, d['value'], d['ms_per_step'], [round(x,1) for x in d['config']['stage_ms_last_step']], d['config']['pcg_iters_last_step'], d['config'].get('pcg_residual_last_step'), d.get('e2e',{}).get('value'))
    for k,v in d['config']['kernels'].items(): print('   ',k, v['launches'], round(v['avg_ms'],4))
except Exception as e:
    print('fail', e); print(open('gpurun_out/bench_n$N.err').read()[-2000:])
PY
done
