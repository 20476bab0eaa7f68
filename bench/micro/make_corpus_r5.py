"""Systems of the reference's own benchmark point (benchmark_cupy.py:28-36: N = 2000, p = 5, r = 5) at d = 50, where nearly
every core of the synthetic fit is rank-deficient: every 7th system of the first two ALS sweeps of the numpy oracle goes to
scratch/corpus.pkl, which bench/micro/vh_corpus.py folds into the corpus of bench/micro/vh_solve.cu.  CPU only."""
import os, pickle, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tt_oracle as O
d, N, r, p = 50, 2000, 5, 5
rng = np.random.RandomState(0)
x = rng.randn(N, d) * np.sqrt(2.0)
y = np.log(0.5 + 0.5 * (x * x).sum(1))
phis = [O.poly_eval(x[:, i], p) for i in range(d)]
np.random.seed(1)
tr = []
O.als(phis, y, n_iter=2, ranks=[1] + [r] * (d - 1) + [1], trace=tr)
out = []
for k, t in enumerate(tr):
    if k % 7 or t["A"].shape[0] != 125: continue
    A = 0.5 * (t["A"] + t["A"].T)
    rk = np.linalg.lstsq(A, t["b"], rcond=None)[2]
    out.append((f"r5 visit {k}", A, t["b"], int(rk)))
os.makedirs(os.path.join(ROOT, "scratch"), exist_ok=True)
pickle.dump(out, open(os.path.join(ROOT, "scratch", "corpus.pkl"), "wb"))
print(len(out), "systems; numerical ranks", [o[3] for o in out])
