# development aid: launch list of the C5 lock-step solve under ncu, summed per kernel, plus one outer iteration in order
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 600 --csv --log-file gpurun_out/c5_launches.csv python tools/run_c5.py --problems 65536 --max-iters 50 > /dev/null 2>&1
python - <<'PY'
import csv,collections,re
rows=[r for r in csv.reader(open('gpurun_out/c5_launches.csv')) if len(r)>5]
h=rows[0]; kn=h.index('Kernel Name'); mv=h.index('Metric Value'); mu=h.index('Metric Unit'); gs=h.index('Grid Size')
agg=collections.defaultdict(lambda:[0,0.0])
seq=[]
for r in rows[1:]:
    name=re.sub(r'\(.*','',r[kn]).replace('void ','').replace('sio::','')
    v=float(r[mv].replace(',','')); u=r[mu]
    v*= {'ns':1e-6,'us':1e-3,'ms':1.0,'s':1e3}.get(u,1.0)
    a=agg[name]; a[0]+=1; a[1]+=v; seq.append((name,round(v*1000,1),r[gs]))
tot=sum(a[1] for a in agg.values())
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f'{k:45s} launches {a[0]:4d} sum ms {a[1]:8.3f}  {a[1]/tot*100:5.1f}%')
print('one round, in order (us):')
i0=[i for i,s in enumerate(seq) if s[0].startswith('k_ls_begin')][3]
for s in seq[i0:i0+26]: print('  ',s)
PY
