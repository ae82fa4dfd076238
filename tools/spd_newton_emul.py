import sys; sys.path.insert(0,'/root/repo/tools'); sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/tmp')
import numpy as np
from dqds_emul import make_pairs, householder_tridiag
exec(open('/tmp/ql_emul.py').read().split("M = make_pairs")[0])
def polish(ds,e2s,x0,simplified,thr2=None):
    P,D=ds.shape; x=x0.astype(np.float64).copy(); ok=np.ones(P,bool)
    for it in range(2):
        p2=np.ones((P,D)); p1=ds[:,:1]-x; q2=np.zeros((P,D)); q1=-np.ones((P,D))
        for k in range(1,D):
            t=ds[:,k:k+1]-x
            pn=t*p1-e2s[:,k-1:k]*p2; qn=t*q1-e2s[:,k-1:k]*q2-p1
            p2,q2,p1,q1=p1,q1,pn,qn
        if it==0: qs=q1; dlt=p1/q1; d1=np.abs(dlt)
        else:
            dlt=p1/(qs if simplified else q1)
            okk = np.abs(dlt)<=np.maximum(1e-3*d1,1e-15)
            if thr2 is not None: okk &= (np.abs(dlt)<=thr2)
            ok&=np.all(okk,axis=1)
        x=x-dlt
    ok&=np.abs(x.sum(1)-1.0)<=1e-13
    return x,ok
M = make_pairs(npairs=400000)
d,e = householder_tridiag(M)
tr=d.sum(1); ds=d/tr[:,None]; es=e/tr[:,None]; e2s=es*es
ref=np.linalg.eigvalsh(M)
x0,its=ql_start_f32(ds,es)
for simp,thr in ((False,None),(True,None),(True,1e-10),(True,1e-11),(True,1e-12)):
    x,ok=polish(ds,e2s,x0,simp,thr)
    lam=np.sort(x*tr[:,None],axis=1); rel=np.max(np.abs(lam-ref)/ref,axis=1)
    s2=np.sum(np.log(lam)**2,1); s2r=np.sum(np.log(ref)**2,1)
    print(simp,thr,'invalid %.2e maxrel(valid) %.2e  max abs err s2 %.2e'%(np.mean(~ok),rel[ok].max(), np.abs(s2-s2r)[ok].max()))

def polish2(ds,e2s,x0,c=5e-17):
    P,D=ds.shape; x=x0.astype(np.float64).copy(); ok=np.ones(P,bool)
    for it in range(2):
        p2=np.ones((P,D)); p1=ds[:,:1]-x; q2=np.zeros((P,D)); q1=-np.ones((P,D))
        for k in range(1,D):
            t=ds[:,k:k+1]-x
            pn=t*p1-e2s[:,k-1:k]*p2
            if it==0: qn=t*q1-e2s[:,k-1:k]*q2-p1
            p2,p1=p1,pn
            if it==0: q2,q1=q1,qn
        if it==0: qs=q1; dlt=p1/q1; d1=np.abs(dlt)
        else:
            dlt=p1/qs
            ok&=np.all(dlt*dlt<=c*d1,axis=1)
        x=x-dlt
    ok&=np.abs(x.sum(1)-1.0)<=1e-13
    return x,ok
for c in (5e-17,2e-16,1e-15):
    x,ok=polish2(ds,e2s,x0,c)
    lam=np.sort(x*tr[:,None],axis=1); rel=np.max(np.abs(lam-ref)/ref,axis=1)
    s2=np.sum(np.log(lam)**2,1); s2r=np.sum(np.log(ref)**2,1)
    print('simplified, est check c=%g'%c,'invalid %.2e maxrel(valid) %.2e  max abs err s2 %.2e'%(np.mean(~ok),rel[ok].max(), np.abs(s2-s2r)[ok].max()))
