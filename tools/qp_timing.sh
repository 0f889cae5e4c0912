# development aid: C5 solve timing (x2) + per-kernel QP time sums under ncu (serialised, comparisons only)
for i in 1 2; do python tools/run_c5.py --problems 65536 --max-iters 50 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['sip']; print('wall',round(s['wall_s'],4),'qp',round(s['qp_s'],4),'sweeps',round(s['oracle_s'],4), s['outer_iterations'])"; done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_sip_qp --csv --log-file gpurun_out/qpx.csv python tools/run_c5.py --problems 65536 --max-iters 50 > /dev/null 2>&1
python - <<'PY'
import csv,collections,re
rows=[r for r in csv.reader(open('gpurun_out/qpx.csv')) if len(r)>5]
h=rows[0]; kn=h.index('Kernel Name'); mv=h.index('Metric Value'); mu=h.index('Metric Unit'); bs=h.index('Block Size')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    m=re.search(r'k_sip_qp<(\d), (\d)>',r[kn])
    key=(m.group(2) if m else r[kn][:30], r[bs])
    v=float(r[mv].replace(',','')); u=r[mu]
    v*= {'ns':1e-6,'us':1e-3,'ms':1.0,'s':1e3,'usecond':1e-3,'nsecond':1e-6,'msecond':1.0}.get(u,1.0)
    a=agg[key]; a[0]+=1; a[1]+=v
for k,a in sorted(agg.items()): print('mode',k[0],'block',k[1],'launches',a[0],'sum ms',round(a[1],2))
PY
