import numpy as np, mpmath as mp, json, sys
from numpy.random import RandomState
from scipy import signal
from oracle import ref_pipeline as rp
mp.mp.dps=50
b,a=rp.butter_highpass(30,16000,5); zi=signal.lfilter_zi(b,a)
pa=mp.polyroots([mp.mpf(float(c)) for c in a], maxsteps=500, extraprec=400)
pb=mp.polyroots([mp.mpf(float(c)) for c in b], maxsteps=2000, extraprec=800)
def split(roots):
    re=sorted([r.real for r in roots if abs(r.imag)<mp.mpf(10)**-30])
    cx=[r for r in roots if r.imag>mp.mpf(10)**-30]
    return re,cx
pr,pc=split(pa); zr,zc=split(pb)
print('poles real',[float(x) for x in pr],'cx',[complex(x) for x in pc]); print('zeros real',[float(x) for x in zr],'cx',[complex(x) for x in zc])
# sections: first-order: real pole with middle real zero; then pairs
def quad_from_pair(r1,r2=None):
    if r2 is None: # complex conj pair
        return (mp.mpf(1), -2*r1.real, r1.real**2+r1.imag**2)
    return (mp.mpf(1), -(r1+r2), r1*r2)
zr_sorted=sorted(zr, key=lambda v: abs(v-1))
sec=[]
# section 1: first order
z1=zr_sorted[0]; p1=pr[0]
sec.append(((mp.mpf(1),-z1,mp.mpf(0)),(mp.mpf(1),-p1,mp.mpf(0))))
zq=[]
rest=zr_sorted[1:]
while len(rest)>=2: zq.append(quad_from_pair(rest[0],rest[1])); rest=rest[2:]
for c in zc: zq.append(quad_from_pair(c))
pq=[quad_from_pair(c) for c in pc]
# pair: pole pair closest to unit circle (largest radius) with zero-quad ... any
for nq,dq in zip(zq,pq): sec.append((nq,dq))
g=mp.mpf(float(b[0]))
sos=np.array([[float(n[0]),float(n[1]),float(n[2]),1.0,float(d[1]),float(d[2])] for n,d in sec])
sos[0,:3]*=float(g)   # gain in the first section
print(sos)
# state-space of DF2T (b,a) and of the cascade, in mp
def ss_df2t(bb,aa):
    n=len(aa)-1
    A=mp.zeros(n,n); B=mp.zeros(n,1); C=mp.zeros(1,n)
    for i in range(n):
        A[i,0]=-aa[i+1]
        if i+1<n: A[i,i+1]=1
        B[i]=bb[i+1]-aa[i+1]*bb[0]
    C[0]=1
    return A,B,C,bb[0]
def series(s1,s2):
    A1,B1,C1,D1=s1; A2,B2,C2,D2=s2
    n1,n2=A1.rows,A2.rows
    A=mp.zeros(n1+n2,n1+n2); B=mp.zeros(n1+n2,1); C=mp.zeros(1,n1+n2)
    for i in range(n1):
        for j in range(n1): A[i,j]=A1[i,j]
        B[i]=B1[i]
    for i in range(n2):
        for j in range(n2): A[n1+i,n1+j]=A2[i,j]
        for j in range(n1): A[n1+i,j]=B2[i]*C1[j]
        B[n1+i]=B2[i]*D1
    for j in range(n1): C[j]=D2*C1[j]
    for j in range(n2): C[n1+j]=C2[j]
    return A,B,C,D2*D1
bq=[mp.mpf(float(c)) for c in b]; aq=[mp.mpf(float(c)) for c in a]
full=ss_df2t(bq,aq)
secs_ss=[]
for k,row in enumerate(sos):
    bb=[mp.mpf(float(v)) for v in row[:3]]; aa=[mp.mpf(float(v)) for v in row[3:]]
    if k==0: bb=bb[:2]; aa=aa[:2]
    secs_ss.append(ss_df2t(bb,aa))
casc=secs_ss[0]
for s in secs_ss[1:]: casc=series(casc,s)
def obs(ss):
    A,B,C,D=ss; n=A.rows; O=mp.zeros(n,n); row=C.copy()
    for k in range(n):
        for j in range(n): O[k,j]=row[j]
        row=row*A
    return O
T=mp.inverse(obs(casc))*obs(full)   # O_c T = O
zic=T*mp.matrix([mp.mpf(float(v)) for v in zi])
zic=np.array([float(v) for v in zic]); print('zi cascade',zic)
# D check
print('D', float(casc[3]), b[0])
zi_sos=np.zeros((3,2)); zi_sos[0,0]=zic[0]; zi_sos[1]=zic[1:3]; zi_sos[2]=zic[3:5]
def ff_casc(x):
    xf=rp.length_fixup(x)
    ext=np.concatenate([2*xf[0]-xf[18:0:-1], xf, 2*xf[-1]-xf[-2:-20:-1]])
    y1,_=signal.sosfilt(sos,ext,zi=zi_sos*ext[0])
    y1=y1.astype(np.float32).astype(np.float64) if F32 else y1
    r=y1[::-1]; y2,_=signal.sosfilt(sos,r,zi=zi_sos*r[0])
    return y2[::-1][18:-18]
LD=np.longdouble
def df2t(x,z0):
    bq=b.astype(LD); aq=a.astype(LD); zs=z0.astype(LD).copy(); y=np.empty(len(x),LD); x=x.astype(LD)
    for n in range(len(x)):
        xn=x[n]; yn=zs[0]+bq[0]*xn
        zs[0]=(zs[1]+xn*bq[1])-yn*aq[1]; zs[1]=(zs[2]+xn*bq[2])-yn*aq[2]; zs[2]=(zs[3]+xn*bq[3])-yn*aq[3]; zs[3]=(zs[4]+xn*bq[4])-yn*aq[4]; zs[4]=xn*bq[5]-yn*aq[5]
        y[n]=yn
    return y
def ff_exact(x):
    xf=rp.length_fixup(x)
    ext=np.concatenate([2*xf[0]-xf[18:0:-1], xf, 2*xf[-1]-xf[-2:-20:-1]])
    y1=df2t(ext,zi*ext[0]); r=y1[::-1].copy(); y2=df2t(r,zi.astype(LD)*r[0])
    return np.asarray(y2[::-1][18:-18],np.float64)
z=np.load('gpurun_out/r2_diag_mel.npz')
par=json.loads(open('scratch/r2_par_scan.json').read().strip().splitlines()[-1])['parity']
cells={c['utt']:c for c in reversed(par['worst_mel_cells'])}
for i in [int(s) for s in sys.argv[1:]] or [4812]:
    pcm=z['pcm%d'%i]; spk,skip,male=z['meta%d'%i]
    x=pcm.astype(np.float64)/32768.0
    def S_of(y):
        prng=RandomState(int(spk)); pos=0
        while pos<skip:
            n=min(skip-pos,1<<22); prng.rand(int(n)); pos+=n
        wav=rp.dither(y,prng); return rp.mel_db_normalize(rp.pySTFT(wav).T)
    yref=signal.filtfilt(b,a,rp.length_fixup(x)); S0=S_of(yref)
    yx=ff_exact(x)
    c=cells[i]; t,bd=c['frame'],c['band']
    for F32 in (False,True):
        yc=ff_casc(x); Sc=S_of(yc)
        print('utt',i,'f32y1' if F32 else 'f64y1','cascade: |y-exact| %.2e |y-scipy| %.2e | dS cell %.2e max dS all %.2e  (exact vs scipy |y| %.2e)'%(np.abs(yc-yx).max(),np.abs(yc-yref).max(),Sc[t,bd]-S0[t,bd],np.abs(Sc-S0).max(),np.abs(yx-yref).max()))
# edge cases: DC, short
rng=np.random.default_rng(11)
tt=np.arange(70000)/16000.0
base=0.3*np.sin(2*np.pi*110*tt)+0.05*rng.standard_normal(tt.shape[0])+0.02
F32=False
for x in [base[:48000], base[:19], base[:257]+0.5, np.zeros(3000), 0.7*np.ones(5000), base[:1000]]:
    print(len(x),'cascade vs scipy %.2e  vs exact %.2e'%(np.abs(ff_casc(x)-signal.filtfilt(b,a,rp.length_fixup(x))).max(), np.abs(ff_casc(x)-ff_exact(x)).max() if len(x)<6000 else -1))
