"""numpy prototype used to choose the eigen-solver of the SVD fallback path: sweeps of the two-sided
cyclic Jacobi method with the round-robin ordering of dense.cu on the graded Gram matrices of bench/micro/eigen_path_check.py,
under different rotation criteria (relative / absolute / sorted diagonal) and after one or two Cholesky LR steps
(B = L^T L).  Result (this container): 13-17 sweeps whatever the criterion, 7-9 after the LR step.  CPU only.

    python bench/micro/jacobi_sweeps.py
"""
import numpy as np, sys
rng=np.random.RandomState(0)
def design(n_l,p,n_r,N,spread):
    L=rng.randn(N,n_l)*np.exp(spread*rng.randn(N,1)); R=rng.randn(N,n_r)*np.exp(spread*rng.randn(N,1))
    x=1.0+0.3*rng.randn(N); phi=np.stack([x**k for k in range(p)],axis=1)
    return np.einsum("sa,sm,sb->samb",L,phi,R).reshape(N,-1)
def rr(k,step,m):
    if k==0: p,q=m-1,step
    else: p,q=(step+k)%(m-1),(step-k+m-1)%(m-1)
    return (p,q) if p<q else (q,p)
def jacobi(A,tolrel=2.2e-16,sort=False,floorfac=1e-3,absfac=0.0):
    A=A.copy(); n=A.shape[0]; eps=2.2e-16
    perm=np.argsort(-np.diag(A)) if sort else np.arange(n)
    A=A[np.ix_(perm,perm)]
    m=(n+1)&~1; hp=m//2
    for sweep in range(40):
        dmax=np.abs(np.diag(A)).max(); floor=max(floorfac*eps*eps*n*dmax, absfac*eps*dmax); rot=0
        for step in range(m-1):
            J=np.eye(n)
            for k in range(hp):
                p,q=rr(k,step,m)
                if q>=n: continue
                app,aqq,apq=A[p,p],A[q,q],A[p,q]
                if abs(apq)>max(tolrel*np.sqrt(abs(app*aqq)),floor):
                    z=(aqq-app)/(2*apq); t=np.sign(z)/(abs(z)+np.sqrt(1+z*z)) if z!=0 else 1.0
                    c=1/np.sqrt(1+t*t); s=c*t
                    J[p,p]=c;J[q,q]=c;J[p,q]=s;J[q,p]=-s; rot+=1
            A=J.T@A@J
        if rot==0: break
    return sweep+1
for (rl,p,rr_,spread) in [(4,4,4,0.0),(4,4,4,1.0),(4,4,4,3.0),(3,3,3,0.5)]:
    V=design(rl,p,rr_,4096,spread); V[:,-1]=V[:,0]+1e-9*V[:,1]
    A=V.T@V; A=0.5*(A+A.T)
    print(rl,p,rr_,spread,"cond %.1e"%np.linalg.cond(A),"base",jacobi(A),"sorted",jacobi(A,sort=True),"tol1e-12",jacobi(A,tolrel=1e-12),"abs 1e-3 eps dmax (no rel)",jacobi(A,tolrel=0.0,absfac=1e-3), "sorted+abs", jacobi(A,tolrel=0.0,absfac=1e-3,sort=True))

print("---- Cholesky-preconditioned (B = L^T L after diagonally pivoted Cholesky) ----")
def pivchol(A, tol):
    A=A.copy(); n=A.shape[0]; L=np.zeros((n,n)); perm=np.arange(n); r=0
    d=np.diag(A).copy()
    for k in range(n):
        j=k+np.argmax(d[k:])
        if d[j] <= tol: break
        # swap k,j
        if j!=k:
            A[[k,j],:]=A[[j,k],:]; A[:,[k,j]]=A[:,[j,k]]; L[[k,j],:]=L[[j,k],:]; perm[[k,j]]=perm[[j,k]]; d[[k,j]]=d[[j,k]]
        L[k,k]=np.sqrt(d[k])
        L[k+1:,k]=(A[k+1:,k]-L[k+1:,:k]@L[k,:k])/L[k,k]
        d[k+1:]-=L[k+1:,k]**2
        r=k+1
    return L[:,:r],perm,r
rng=np.random.RandomState(0)
for (rl,p,rr_,spread) in [(4,4,4,0.0),(4,4,4,1.0),(4,4,4,3.0),(3,3,3,0.5)]:
    V=design(rl,p,rr_,4096,spread); V[:,-1]=V[:,0]+1e-9*V[:,1]
    A=V.T@V; A=0.5*(A+A.T); n=A.shape[0]
    eps=2.2e-16
    L,perm,r=pivchol(A, 1e-3*eps*eps*n*np.diag(A).max())
    B=L.T@L
    s1=jacobi(B)
    L2,perm2,r2=pivchol(B, 0.0); B2=L2.T@L2
    s2=jacobi(B2)
    ev=np.sort(np.linalg.eigvalsh(A))[::-1]; evB=np.sort(np.linalg.eigvalsh(B))[::-1]
    print(rl,p,rr_,spread,"rank",r,"sweeps on B",s1,"on B2 (two LR steps)",s2,"eig rel err top/bottom", abs(evB[0]-ev[0])/ev[0], abs(evB[r-1]-ev[r-1])/abs(ev[r-1]))
