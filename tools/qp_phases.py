"""Per-phase breakdown of k_sip_qp from an ncu report (development aid; how profiles/r02_qp_phases.md was made).

    ncu --set full --clock-control none --import-source on -k regex:k_sip_qp -s 36 -c 2 -o gpurun_out/qp \
        python tools/run_c5.py --problems 65536 --max-iters 50
    python tools/qp_phases.py gpurun_out/qp.ncu-rep ELi1E [launch index]

Maps every SASS instruction of the chosen instantiation (ELi0E full, ELi1E narrow, ELi2E wide) to a source line with
`nvdisasm -g`, classes the lines into solver phases by the markers found in sio_qp.cu, and sums ncu's per-instruction
executed counts and warp-state samples (source page) per phase.  Needs the in-tree libsio_b200.so the report was taken with.
"""
import re,csv,collections,subprocess,sys
rep=sys.argv[1]; mode=sys.argv[2]   # e.g. ELi1E
import os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so=os.path.join(ROOT,'semiinfiniteoptimization_b200/lib/libsio_b200.so')
import os,tempfile
d=tempfile.mkdtemp()
subprocess.run(f'cd {d} && cuobjdump -xelf sio_qp {so} >/dev/null && nvdisasm -g -c sio_qp.sm_100a.cubin > dis.txt',shell=True,check=True)
lines=open(f'{d}/dis.txt').read().split('\n')
start=[i for i,l in enumerate(lines) if l.startswith('.text._ZN3sio8k_sip_qpILi6'+mode)][0]
src=open(os.path.join(ROOT,'semiinfiniteoptimization_b200/csrc/sio_qp.cu')).read().split('\n')
def find(pat): return [i+1 for i,l in enumerate(src) if pat in l][0]
L_inv0=find('struct SpdInverse'); L_inv1=find('__device__ __forceinline__ void apply')
L_lu0=find('__device__ bool warp_lu_solve'); L_lu1=find('// one lane\'s share of a problem')
L_setrho0=find('auto set_rho = [&]()'); L_setrho1=find('    for (int s = 0; s < NR; ++s) sv[s] = 0.0;')
L_chk=find('// residuals (qp.py solve_qp'); L_cert=find('// primal infeasibility certificate'); L_adapt=find('if (it % adapt_every == 0')
L_pol=find('// ---- polish on the active set'); L_admm_end=find('// status codes written per problem')
L_setup0=find('__device__ __forceinline__ void solve_rows'); 
def phase(p):
    if not p or p[0]!='sio_qp.cu': return None
    n=p[1]
    if L_inv0<=n<L_inv1: return 'set_rho'
    if L_lu0<=n<L_lu1: return 'polish_lu'
    if L_setrho0<=n<L_setrho1: return 'set_rho'
    if L_setrho1<=n<L_chk: return 'hot'
    if L_chk<=n<L_cert: return 'check'
    if L_cert<=n<L_adapt: return 'certificate'
    if L_adapt<=n<L_pol: return 'adapt'
    if L_pol<=n<L_admm_end: return 'polish_rest'
    if n>=L_setup0: return 'setup'
    return None
cur=None; per=[]
for l in lines[start+1:]:
    if l.startswith('//-----') : break
    m=re.match(r'\s*//## File "([^"]+)", line (\d+)',l)
    if m: cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S',l): per.append(cur)
ph=[phase(p) for p in per]
n=len(ph); left=[None]*n
last=None;dd=10**9
for i in range(n):
    if ph[i]: last=ph[i]; dd=0
    else: dd+=1
    left[i]=(last,dd)
last=None;dd=10**9; out=[None]*n
for i in range(n-1,-1,-1):
    if ph[i]: last=ph[i]; dd=0
    else: dd+=1
    l=left[i]
    out[i]= ph[i] or (l[0] if l[1]<=dd else last)
kid=sys.argv[3] if len(sys.argv)>3 else '0'
txt=subprocess.run(f'ncu -i {rep} --page source --csv --print-source sass --launch-skip {kid} --launch-count 1',shell=True,capture_output=True,text=True).stdout
rows=list(csv.reader(txt.split('\n')))
R=[]; h=None
for r in rows:
    if not r: continue
    if r[0]=='Kernel Name':
        if R: break
        print(r[1][:60]); continue
    if r[0]=='Address': h=r; continue
    if h: R.append(r)
assert len(R)==n,(len(R),n)
ie=h.index('Instructions Executed'); sa=h.index('# Samples'); sn=h.index('stall_no_inst')
tot=sum(int(r[sa]) for r in R); tote=sum(int(r[ie]) for r in R)
agg=collections.defaultdict(lambda:[0,0,0,0,0])
for o,r in zip(out,R):
    e=int(r[ie]); a=agg[o]; a[0]+=1; a[1]+=e; a[2]+=int(r[sa]); a[3]+=int(r[sn]); a[4]+=(1 if e>0 else 0)
print('total instructions',tote,'samples',tot,'max exec per instr',max(int(r[ie]) for r in R))
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][2]):
    print(f'{str(k):12s} static {a[0]:6d} (executed {a[4]:5d})  exec {a[1]/tote*100:5.1f}%  time {a[2]/tot*100:5.1f}%  of which no_inst {a[3]/max(1,a[2])*100:5.1f}%')
