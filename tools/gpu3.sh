ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_b.log 2>&1
python - <<'PY'
import csv, collections, re
rows=[]
with open('gpurun_out/launches.csv') as f:
    lines=[l for l in f if l.startswith('"')]
r=csv.DictReader(lines)
agg=collections.defaultdict(lambda:[0,0.0])
for row in r:
    if row.get('Metric Name')!='gpu__time_duration.sum': continue
    name=row['Kernel Name']; name=re.sub(r'\(.*','',name)[:70]
    v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
    if u=='ns': v/=1e6
    elif u=='us': v/=1e3
    elif u=='s': v*=1e3
    agg[name][0]+=1; agg[name][1]+=v
tot=sum(v[1] for v in agg.values())
print("total ms",tot)
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:40]:
    print("%-72s %5d %9.3f %8.4f %5.1f%%"%(k,v[0],v[1],v[1]/v[0],100*v[1]/tot))
PY
