"""How the fan-plane kernel's lanes fill up (csrc/ax_fan.cu: greedy packing of the z ranges of a column's runs into
the 32 lanes of a pass, per epoch) at C2, for a few angles and columns and several epoch lengths: passes per column,
pass-steps per column (the kernel's per-step cost is per pass) and the share of busy lanes.  Host-side NumPy, no GPU.

    python tools/fan_pack_sim.py
"""
import sys, numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tools'))
from fan_event_model import frame
from physics_arx_b200 import phantom
cfg = phantom.config_c2(); geo = cfg['geo']
def sim(theta, u, EP, CW=8):
    F = frame(geo, theta); nv=F['nv']
    sx,sy,sz,vz = F['sx'],F['sy'],F['sz'],F['vz']
    dx,dy = F['ox']+u*F['ux']-sx, F['oy']+u*F['uy']-sy
    dz0 = F['oz']-sz
    rows=np.arange(nv)
    nrow=np.ceil(np.sqrt(dx*dx+dy*dy+(dz0+rows*vz)**2)/F['acc']).astype(int)
    starts=np.flatnonzero(np.r_[True,nrow[1:]!=nrow[:-1]]); ends=np.r_[starts[1:],nv]
    def slab(s,d,lo,hi,t0,t1):
        if d==0: return (t0,t1) if lo<=s<hi else (1,0)
        ta,tb=(lo-s)/d,(hi-s)/d
        return max(t0,min(ta,tb)),min(t1,max(ta,tb))
    l0,l1=slab(sx,dx,0,F['hix'],0,1); l0,l1=slab(sy,dy,0,F['hiy'],l0,l1)
    if l0>l1: return 0,0,0
    Emin=int(np.floor(l0*nrow.min())); Emax=int(np.ceil(l1*nrow.max()))
    kmin,kmax=3,132
    passes=0; steps=0; lanes_tot=0
    for e0 in range(Emin,Emax+1,EP):
        e1=min(e0+EP-1,Emax)
        used=0; nseg=0; npass=0; lanes=[]
        for ra,rb in zip(starts,ends):
            n=nrow[ra]; I0=int(np.ceil(l0*n)); I1=int(np.floor(l1*n)); ja,jb=max(I0,e0),min(I1,e1)
            if ja>jb: continue
            cur=ra
            while cur<rb:
                s=lambda r:(dz0+r*vz)/n
                zlo=sz+min(ja*s(cur),jb*s(cur)); ka=max(kmin,int(np.floor(zlo)));
                if ka>kmax: cur=rb; break
                qa=ka>>2; avail=32-used
                r2=rb
                def nq(r2):
                    zhi=sz+max(ja*s(r2-1),jb*s(r2-1)); kb=max(min(kmax,int(np.floor(zhi))+1),ka); return (kb>>2)-qa+1
                while nq(r2)>avail and r2>cur+1: r2-=1
                q=nq(r2)
                if q>avail or nseg>=8:
                    lanes.append(used); used=0; nseg=0; continue
                used+=q; nseg+=1; cur=r2
        if used: lanes.append(used)
        nwin=(e1-e0+CW)//CW
        passes+=len(lanes); steps+=len(lanes)*nwin*CW; lanes_tot+=sum(lanes)*nwin*CW
    return passes, steps, lanes_tot
for EP in (64,128,256,512):
    tot=np.zeros(3)
    for th in (0.1,0.7,1.3,2.9,4.0):
        for u in (10,200,400,600,800,1000):
            tot+=np.array(sim(th,u,EP))
    print(EP, 'passes',tot[0]/30,'pass-steps/col',tot[1]/30,'lane util',tot[2]/tot[1]/32)
