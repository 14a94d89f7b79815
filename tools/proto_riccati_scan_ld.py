import numpy as np, sys
from proto_riccati_scan import *
def seq_ld(A,B,Hu,Hx,PN):
    ld=np.longdouble
    N=len(A); P=[None]*(N+1); P[N]=PN.astype(ld)
    for k in range(N-1,0,-1):
        Pn=P[k+1]; Ak=A[k].astype(ld);Bk=B[k].astype(ld)
        R=np.diag(Hu[k].astype(ld))+Bk.T@Pn@Bk; S=Bk.T@Pn@Ak
        # m=1
        K=-S/R[0,0]; P[k]=np.diag(Hx[k].astype(ld))+Ak.T@Pn@Ak+S.T@K
        P[k]=(P[k]+P[k].T)/2
    return P
rng=np.random.default_rng(1)
for hi in [4,7,10]:
  for n,N in [(2,30),(3,30),(2,100)]:
    ws=wp=0
    for trial in range(200):
        dt=0.02
        if n==2: Ac=np.array([[1,dt],[0,1.]]); Bc=np.array([[dt*dt/2],[dt]])*rng.uniform(.5,20)
        else: Ac=np.array([[1,dt,dt*dt/2],[0,1,dt],[0,0,1.]]); Bc=np.array([[dt**3/6],[dt*dt/2],[dt]])
        A=[Ac]*N;B=[Bc]*N
        Hu=[10**rng.uniform(-3,hi,size=1) for _ in range(N)]
        Hx=[10**rng.uniform(-3,hi,size=n)*(rng.random(n)<0.7)+10**rng.uniform(-2,2,size=n) for _ in range(N)]
        PN=np.diag(10**rng.uniform(-2,3,size=n))
        Pt=seq_ld(A,B,Hu,Hx,PN);Ps=seq(A,B,Hu,Hx,PN);Pp=par(A,B,Hu,Hx,PN)
        for k in range(1,N):
            sc=float(np.abs(Pt[k]).max())
            ws=max(ws,float(np.abs(Ps[k]-Pt[k]).max())/sc); wp=max(wp,float(np.abs(Pp[k]-Pt[k]).max())/sc)
    print(hi,n,N,"seq err",ws,"par err",wp)
