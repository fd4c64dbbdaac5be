timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "N=2 rc $?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_n2.json') if l.startswith('{')][-1])
print(2, d['value'], d['ms_per_step'], [round(x,1) for x in d['config']['stage_ms_last_step']], d.get('e2e',{}).get('value'))
PY
