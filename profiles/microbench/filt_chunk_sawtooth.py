import numpy as np, mpmath as mp, json, sys
from numpy.random import RandomState
from scipy import signal
from oracle import ref_pipeline as rp
z=np.load('gpurun_out/r2_diag_mel.npz')
par=json.loads(open('scratch/r2_par_scan.json').read().strip().splitlines()[-1])['parity']
cells={c['utt']:c for c in reversed(par['worst_mel_cells'])}
b,a=rp.butter_highpass(30,16000,5); zi=signal.lfilter_zi(b,a)
LD=np.longdouble
def df2t(bq,aq,x,z0,dt):
    bq=bq.astype(dt); aq=aq.astype(dt); zs=z0.astype(dt).copy(); y=np.empty(len(x),dt)
    x=x.astype(dt)
    for n in range(len(x)):
        xn=x[n]; yn=zs[0]+bq[0]*xn
        zs[0]=(zs[1]+xn*bq[1])-yn*aq[1]; zs[1]=(zs[2]+xn*bq[2])-yn*aq[2]; zs[2]=(zs[3]+xn*bq[3])-yn*aq[3]; zs[3]=(zs[4]+xn*bq[4])-yn*aq[4]; zs[4]=xn*bq[5]-yn*aq[5]
        y[n]=yn
    return y,zs
def df2t_states(bq,aq,x,z0,dt,every):
    """exact-ish run in dt, returning output and the state at every `every` samples"""
    bq=bq.astype(dt); aq=aq.astype(dt); zs=z0.astype(dt).copy(); y=np.empty(len(x),dt); st=[]
    x=x.astype(dt)
    for n in range(len(x)):
        if n%every==0: st.append(zs.copy())
        xn=x[n]; yn=zs[0]+bq[0]*xn
        zs[0]=(zs[1]+xn*bq[1])-yn*aq[1]; zs[1]=(zs[2]+xn*bq[2])-yn*aq[2]; zs[2]=(zs[3]+xn*bq[3])-yn*aq[3]; zs[3]=(zs[4]+xn*bq[4])-yn*aq[4]; zs[4]=xn*bq[5]-yn*aq[5]
        y[n]=yn
    return y,st
def chunked(x,z0,chunk):
    """fp64 scipy-order recurrence inside chunks, entry states from the long-double run"""
    yl,st=df2t_states(b,a,x,z0,LD,chunk)
    y=np.empty(len(x))
    for c,s in enumerate(st):
        seg=x[c*chunk:(c+1)*chunk]
        y[c*chunk:(c+1)*chunk],_=signal.lfilter(b,a,seg,zi=s.astype(np.float64))
    return y, yl.astype(np.float64)
def filtfilt_variant(x, chunk):
    xf=rp.length_fixup(x)
    ext=np.concatenate([2*xf[0]-xf[18:0:-1], xf, 2*xf[-1]-xf[-2:-20:-1]])
    y1,y1x=chunked(ext, zi*ext[0], chunk)
    r=y1[::-1].copy(); y2,y2x=chunked(r, zi*r[0], chunk)
    rx=y1x[::-1].copy(); y2xx,_=df2t(b,a,rx,zi*rx[0],LD)
    return y2[::-1][18:-18], np.asarray(y2xx[::-1][18:-18],np.float64)
which=[int(s) for s in sys.argv[1:]] or [4812]
for i in which:
    pcm=z['pcm%d'%i]; spk,skip,male=z['meta%d'%i]
    x=pcm.astype(np.float64)/32768.0
    def S_of(y):
        prng=RandomState(int(spk)); pos=0
        while pos<skip:
            n=min(skip-pos,1<<22); prng.rand(int(n)); pos+=n
        wav=rp.dither(y,prng); return rp.mel_db_normalize(rp.pySTFT(wav).T)
    yref=signal.filtfilt(b,a,rp.length_fixup(x)); S0=S_of(yref)
    c=cells[i]; t,bd=c['frame'],c['band']
    print('utt',i,'cell',(t,bd),'GPU scan dS %.2e'%(z['mel_prod%d'%i][t,bd]-S0[t,bd]))
    for chunk in (256,64):
        yc,yx=filtfilt_variant(x,chunk)
        Sc=S_of(yc); Sx=S_of(yx)
        print('  chunk %3d: |y-scipy| max %.2e  dS at cell %.2e  max dS band0-1 %.2e | exact(longdouble): |y-scipy| %.2e dS cell %.2e max band0-1 %.2e'%(chunk,np.abs(yc-yref).max(),Sc[t,bd]-S0[t,bd],np.abs(Sc-S0)[:,:2].max(),np.abs(yx-yref).max(),Sx[t,bd]-S0[t,bd],np.abs(Sx-S0)[:,:2].max()))
