import sys; sys.path.insert(0,'/root/repo/tools'); sys.path.insert(0,'/root/repo')
import numpy as np
from dqds_emul import make_pairs, householder_tridiag, newton_polish
f32=np.float32

def ql_start_f32(ds, es, schedule=None, tolshift=16, maxit=6):
    """emulates ql_tri_eigvals_start_f32. schedule: list of fixed iteration counts per l, or None for convergence test.
    returns eigen starts and iteration counts [P, D-1]"""
    P,D = ds.shape
    d = ds.astype(f32).copy(); e = np.zeros((P,D),dtype=f32); e[:,:D-1]=es.astype(f32)
    its = np.zeros((P,D-1),dtype=np.int32)
    with np.errstate(all='ignore'):
        for l in range(D-1):
            nit = maxit if schedule is None else schedule[l]
            for it in range(nit):
                if schedule is None:
                    he = np.abs(e[:,l]); hm = np.maximum(np.abs(d[:,l]),np.abs(d[:,l+1]))
                    act = ~(he*f32(2.0**tolshift) <= hm)
                    if not act.any(): break
                else:
                    act = np.ones(P,dtype=bool)
                its[act,l]+=1
                dl = f32(0.5)*(d[:,l+1]-d[:,l])
                e2 = e[:,l]*e[:,l]
                hh = dl*dl+e2
                g = d[:,D-1]-d[:,l] + e2/(dl+np.copysign(np.sqrt(hh),dl))
                s=np.ones(P,dtype=f32); c=np.ones(P,dtype=f32); p=np.zeros(P,dtype=f32)
                dn=d.copy(); en=e.copy()
                for i in range(D-2,l-1,-1):
                    f=s*en[:,i]; b=c*en[:,i]
                    h2=f*f+g*g
                    rr=f32(1)/np.sqrt(h2)
                    en[:,i+1]=h2*rr
                    s=f*rr; c=g*rr
                    g=dn[:,i+1]-p
                    r2=(dn[:,i]-g)*s+f32(2)*c*b
                    p=s*r2
                    dn[:,i+1]=g+p
                    g=c*r2-b
                dn[:,l]-=p; en[:,l]=g; en[:,D-1]=0
                d=np.where(act[:,None],dn,d); e=np.where(act[:,None],en,e)
    return d, its

M = make_pairs(npairs=200000)
d,e = householder_tridiag(M)
tr=d.sum(1); ds=d/tr[:,None]; es=e/tr[:,None]; e2s=es*es
ref=np.linalg.eigvalsh(M)
def evaluate(label, **kw):
    x0,its = ql_start_f32(ds,es,**kw)
    x,ok = newton_polish(ds,e2s,x0)
    lam=np.sort(x*tr[:,None],axis=1)
    rel=np.max(np.abs(lam-ref)/ref,axis=1)
    P=len(M); nw=P//32
    w = its[:nw*32].reshape(nw,32,-1).max(axis=1)
    D=6
    steps_lane=(its*(D-1-np.arange(D-1))[None,:]).sum(1).mean()
    steps_warp=(w*(D-1-np.arange(D-1))[None,:]).sum(1).mean()
    print(f'{label}: invalid {np.mean(~ok):.2e} maxrel(valid) {rel[ok].max():.1e} iters/lane {its.mean(0).round(2)} iters/warp {w.mean(0).round(2)} steps lane {steps_lane:.1f} warp {steps_warp:.1f}')
evaluate('conv 2^-16')
evaluate('conv 2^-11', tolshift=11)
for sch in ([3,3,2,2,2],[3,2,2,2,2],[3,2,2,2,1],[2,2,2,2,2],[3,3,3,2,2],[4,3,2,2,2],[3,3,3,3,2],[4,3,3,2,2]):
    evaluate(f'fixed {sch}', schedule=sch)
