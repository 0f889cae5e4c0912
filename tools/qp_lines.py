"""Per-source-line executed instructions / samples of k_sip_qp from an ncu report (development aid, see qp_phases.py).

    python tools/qp_lines.py gpurun_out/qp.ncu-rep ELi1E <first line> <last line>
"""
import re,csv,collections,subprocess,sys,tempfile
rep=sys.argv[1]; mode=sys.argv[2]; lo=int(sys.argv[3]); hi=int(sys.argv[4])
import os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so=os.path.join(ROOT,'semiinfiniteoptimization_b200/lib/libsio_b200.so')
d=tempfile.mkdtemp()
subprocess.run(f'cd {d} && cuobjdump -xelf sio_qp {so} >/dev/null && nvdisasm -g -c sio_qp.sm_100a.cubin > dis.txt',shell=True,check=True)
lines=open(f'{d}/dis.txt').read().split('\n')
start=[i for i,l in enumerate(lines) if l.startswith('.text._ZN3sio8k_sip_qpILi6'+mode)][0]
cur=None; per=[]; txt_i=[]
for l in lines[start+1:]:
    if l.startswith('//-----') : break
    m=re.match(r'\s*//## File "([^"]+)", line (\d+)',l)
    if m: cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    m=re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(\S.*)',l)
    if m: per.append(cur); txt_i.append(m.group(1))
txt=subprocess.run(f'ncu -i {rep} --page source --csv --print-source sass --launch-skip 0 --launch-count 1',shell=True,capture_output=True,text=True).stdout
rows=list(csv.reader(txt.split('\n')))
R=[]; h=None
for r in rows:
    if not r: continue
    if r[0]=='Kernel Name':
        if R: break
        continue
    if r[0]=='Address': h=r; continue
    if h: R.append(r)
ie=h.index('Instructions Executed'); sa=h.index('# Samples'); sn=h.index('stall_no_inst')
tot=sum(int(r[sa]) for r in R); tote=sum(int(r[ie]) for r in R)
src=open(os.path.join(ROOT,'semiinfiniteoptimization_b200/csrc/sio_qp.cu')).read().split('\n')
agg=collections.defaultdict(lambda:[0,0,0,0])
for p,r in zip(per,R):
    if p and p[0]=='sio_qp.cu' and lo<=p[1]<=hi:
        a=agg[p[1]]; a[0]+=1; a[1]+=int(r[ie]); a[2]+=int(r[sa]); a[3]+=int(r[sn])
for k in sorted(agg):
    a=agg[k]
    print(f'{k:4d} static {a[0]:4d} exec {a[1]/tote*100:5.2f}% time {a[2]/tot*100:5.2f}% noinst {a[3]/tot*100:5.2f}% | {src[k-1].strip()[:90]}')
