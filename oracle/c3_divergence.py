"""TEST INFRASTRUCTURE (oracle side, CPU only) — evidence for DESIGN.md §9: the reference ALGORITHM does not converge on
BASELINE.json's configs[2] problem (HJB, X0 = 0, sigma = sqrt(2), raw monomials, random re-initialisation of the cores at
every fit, als.py:36-37), independently of any GPU code.  Runs the numpy restatement of the reference's explicit scheme
(oracle/tt_oracle.py, pinned against the reference by tests/test_oracle_golden.py) at reduced N:

    python oracle/c3_divergence.py 100 1024 50 hjb 4 4 ALS      # d, N, steps, model, rank, p, optimiser
    python oracle/c3_divergence.py 20 2048 10 hjb 4 4 ALS
    python oracle/c3_divergence.py 100 1024 10 hjb 4 4 SALSA

Observed (this container): the terminal fit leaves a relative residual of 0.97, grad V is O(1e6), h = -|Z|^2/2 squares it
at every step, numpy overflows after 5 steps (ALS) or LAPACK's SVD fails to converge (SALSA) — the reference would raise.
The CUDA path follows the same trajectory (bench/diag_solve.py): non-finite Y from step 45 of 50 at N = 2^18.
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, oracle.tt_oracle as O
d,N,steps,model_name=int(sys.argv[1]),int(sys.argv[2]),int(sys.argv[3]),sys.argv[4]
r=int(sys.argv[5]); p=int(sys.argv[6]); opt=sys.argv[7]
np.random.seed(0)
if model_name=="hjb":
    model=O.HJB(1.0,np.sqrt(2.0)); X0=np.zeros((N,d)); ref=4.5901
else:
    model=O.BlackScholes(1.0,0.05,0.4); x0=np.array([1.0,0.5]*(d//2)); X0=np.tile(x0,(N,1)); ref=np.exp(0.21)*np.sum(x0**2)
X,xi=O.generate_trajectories(X0,steps,model)
t0=time.time()
Y,c=O.explicit_scheme(X,model,p=p,ranks=[1]+[r]*(d-1)+[1],n_iter=10,optimizer_name=opt,salsa_kwargs=dict(max_rank=5) if opt=="SALSA" else None)
print(model_name,d,N,opt,"Y0",Y[:,0].mean(),"ref",ref,"nan steps",[n for n in range(steps+1) if not np.isfinite(Y[:,n]).all()][:3],"t",time.time()-t0)
print("Y means",[float(Y[:,n].mean()) for n in range(0,steps+1,5)])
