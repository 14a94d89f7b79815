import numpy as np
rng=np.random.default_rng(0)
def seq(A,B,Hu,Hx,PN):
    N=len(A); P=[None]*(N+1); P[N]=PN
    for k in range(N-1,0,-1):
        Pn=P[k+1]; R=np.diag(Hu[k])+B[k].T@Pn@B[k]; S=B[k].T@Pn@A[k]
        K=-np.linalg.solve(R,S); P[k]=np.diag(Hx[k])+A[k].T@Pn@A[k]+S.T@K
        P[k]=(P[k]+P[k].T)/2
    return P
def comb(E1,E2):  # E1 outer (earlier stage), E2 inner
    A1,C1,J1=E1;A2,C2,J2=E2
    n=A1.shape[0]
    M=np.linalg.inv(np.eye(n)+C1@J2)
    A12=A2@M@A1
    C12=A2@M@C1@A2.T+C2; C12=(C12+C12.T)/2
    J12=A1.T@J2@M@A1+J1; J12=(J12+J12.T)/2
    return (A12,C12,J12)
def par(A,B,Hu,Hx,PN,slots=16):
    N=len(A); n=A[0].shape[0]
    # visit order t=0..N-2 -> k=N-1-t
    q=-(-N//slots)
    I=(np.eye(n),np.zeros((n,n)),np.zeros((n,n)))
    el=lambda k:(A[k],B[k]@np.diag(1/Hu[k])@B[k].T,np.diag(Hx[k]))
    comp=[]
    for sl in range(slots):
        E=(np.zeros((n,n)),np.zeros((n,n)),PN) if sl==0 else I
        for j in range(q):
            t=sl*q+j
            if t>N-2: break
            E=comb(el(N-1-t),E)
        comp.append(E)
    d=1
    while d<slots:
        new=list(comp)
        for sl in range(d,slots): new[sl]=comb(comp[sl],comp[sl-d])
        comp=new; d*=2
    P=[None]*(N+1); P[N]=PN
    for sl in range(slots):
        Pn=PN if sl==0 else comp[sl-1][2]
        for j in range(q):
            t=sl*q+j
            if t>N-2: break
            k=N-1-t
            R=np.diag(Hu[k])+B[k].T@Pn@B[k]; S=B[k].T@Pn@A[k]
            K=-np.linalg.solve(R,S); Pk=np.diag(Hx[k])+A[k].T@Pn@A[k]+S.T@K
            Pk=(Pk+Pk.T)/2; P[k]=Pk; Pn=Pk
    return P
for n,N in [(2,30),(3,30),(2,100),(3,100)]:
    worst=0
    for trial in range(300):
        dt=0.02
        if n==2: Ac=np.array([[1,dt],[0,1.]]); Bc=np.array([[dt*dt/2],[dt]])*rng.uniform(.5,20)
        else: Ac=np.array([[1,dt,dt*dt/2],[0,1,dt],[0,0,1.]]); Bc=np.array([[dt**3/6],[dt*dt/2],[dt]])
        A=[Ac]*N;B=[Bc]*N
        Hu=[10**rng.uniform(-3,10,size=1) for _ in range(N)]
        Hx=[10**rng.uniform(-3,10,size=n)*(rng.random(n)<0.7)+10**rng.uniform(-2,2,size=n) for _ in range(N)]
        PN=np.diag(10**rng.uniform(-2,3,size=n))
        Ps=seq(A,B,Hu,Hx,PN);Pp=par(A,B,Hu,Hx,PN)
        for k in range(1,N):
            e=np.abs(Ps[k]-Pp[k]).max()/np.abs(Ps[k]).max()
            worst=max(worst,e)
    print(n,N,worst)
