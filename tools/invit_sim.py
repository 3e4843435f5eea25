#!/usr/bin/env python
"""CPU model (NumPy, vectorised over the eigenvalues) of the inverse-iteration stage of csrc/trieig.cu on the tridiagonal
matrix of the paper-4 stamp (dumped by tools/eig_p4_debug.py into gpurun_out/tridiag_p4.npz): conditioning of the Gram
matrix of the Cholesky-QR stage and residuals, as a function of the minimum spacing enforced between shifts (in
ulp(|T|)) and of the number of solve + orthonormalise stages.

    python tools/invit_sim.py <min spacing / ulp|T|> <stages> [seed]

Findings: spreading close shifts by multiples of ulp(|T|) makes the vectors dependent (spacing 10:
singular Gram matrix; 1: min eigenvalue 1e-10; 0.5: 4e-8); with the computed eigenvalues themselves as shifts the
vectors are orthogonal to 1e-15 before any orthogonalisation and the residual is 13 ulp(|T|)."""
import numpy as np, sys, time, scipy.linalg as sl
import os
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out', 'tridiag_p4.npz'))
d, e, lt = z['d'], z['e'], z['lt']
n = d.size
lam = np.sort(lt)
tn = max(abs(lam[0]), abs(lam[-1])); ulp = 2.220446049250313e-16*tn
dmin = float(sys.argv[1])*ulp; p=int(sys.argv[2]); seed=int(sys.argv[3]) if len(sys.argv)>3 else 1
xs = lam.copy()
for j in range(1,n): xs[j]=max(lam[j], xs[j-1]+dmin)
print("max push (ulp)", ((xs-lam)/ulp).max())
N=n
def factor(xs):
    u0=np.zeros((n,N)); u1=np.zeros((n,N)); u2=np.zeros((n,N)); mm=np.zeros((n,N)); pv=np.zeros((n,N),bool)
    a=d[0]-xs; b=np.full(N,e[0]); c=np.zeros(N); tiny=ulp
    for i in range(n-1):
        sub=e[i]; dn=d[i+1]-xs; en=e[i+1] if i+2<n else 0.0
        nosw=np.abs(a)>=abs(sub)
        a=np.where(nosw & (a==0),tiny,a)
        m=np.where(nosw, sub/a, a/sub)
        u0[i]=np.where(nosw,a,sub); u1[i]=np.where(nosw,b,dn); u2[i]=np.where(nosw,c,en); pv[i]=~nosw; mm[i]=m
        a2=np.where(nosw, dn-m*b, b-m*dn); b2=np.where(nosw, en-m*c, c-m*en)
        a,b,c=a2,b2,np.zeros(N)
    a=np.where(a==0,tiny,a); u0[n-1]=a
    return u0,u1,u2,mm,pv
u0,u1,u2,mm,pv=factor(xs)
def solve(x):
    x=x.copy()
    for i in range(n-1):
        xi=x[i].copy(); xn=x[i+1].copy(); sw=pv[i]; m=mm[i]
        x[i]=np.where(sw,xn,xi); x[i+1]=np.where(sw, xi-m*xn, xn-m*xi)
    x1=np.zeros(N); x2=np.zeros(N)
    with np.errstate(over='ignore',invalid='ignore'):
        for i in range(n-1,-1,-1):
            v=(x[i]-u1[i]*x1-u2[i]*x2)/u0[i]
            big=np.abs(v)>1e150
            if big.any():
                x[:,big]*=1e-150; v=np.where(big,v*1e-150,v); x1=np.where(big,x1*1e-150,x1)
            x[i]=v; x2=x1; x1=v
    x/= np.abs(x).max(axis=0)
    x/=np.sqrt((x*x).sum(axis=0))
    return x
T=np.diag(d)+np.diag(e[:n-1],1)+np.diag(e[:n-1],-1)
def cholqr(Z, tag):
    G=Z@Z.T; w=np.linalg.eigvalsh(G)
    print(f"  [{tag}] G eig min {w.min():.3e} max {w.max():.3e} #<1e-10 {(w<1e-10).sum()}")
    L=np.linalg.cholesky(G)
    return sl.solve_triangular(L,Z,lower=True)
def report(Q, tag):
    R=Q@T-lam[:,None]*Q
    print(f"{tag}: orth {np.abs(Q@Q.T-np.eye(n)).max():.2e} resid max {np.abs(R).max()/ulp:.1f} ulp|T|, row-norm max {np.sqrt((R*R).sum(1)).max()/ulp:.1f} ulp")
rng=np.random.default_rng(seed)
x=rng.uniform(-1,1,(n,N))
x=solve(x)
Q=cholqr(x.T,"s1")
for k in range(p-1):
    x=solve(Q.T); Q=cholqr(x.T,f"s{k+2}")
Q=cholqr(Q,"final"); report(Q,f"dmin={sys.argv[1]}ulp stages={p}")
