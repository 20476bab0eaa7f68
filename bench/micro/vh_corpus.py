"""Writes the corpus for bench/micro/vh_solve.cu: normal-equation systems of real ALS fits (scratch/make_corpus.py) plus
synthetic graded / rank-deficient Gram matrices, each with numpy.linalg.lstsq's solution and numerical rank."""
import os, pickle, struct, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
out = []
pk = os.path.join(ROOT, "scratch", "corpus.pkl")
if os.path.exists(pk):
    for tag, A, b, rk in pickle.load(open(pk, "rb"))[::3]:
        out.append((0.5 * (A + A.T), b))
rng = np.random.RandomState(0)
def design(n_l, p, n_r, N, spread):
    L = rng.randn(N, n_l) * np.exp(spread * rng.randn(N, 1)); R = rng.randn(N, n_r) * np.exp(spread * rng.randn(N, 1))
    x = 1.0 + 0.3 * rng.randn(N)
    return np.einsum("sa,sm,sb->samb", L, np.stack([x**k for k in range(p)], axis=1), R).reshape(N, -1)
for (rl, p, rr, spread) in [(4, 4, 4, 0.0), (4, 4, 4, 1.0), (4, 4, 4, 3.0), (2, 3, 2, 1.0), (4, 4, 1, 2.0), (5, 4, 5, 1.0), (3, 3, 3, 0.5), (5, 5, 5, 0.5), (6, 4, 6, 0.3)]:
    V = design(rl, p, rr, 4096, spread)
    V[:, -1] = V[:, 0] + 1e-9 * V[:, 1]
    y = np.log(1 + np.sum(V[:, :3] ** 2, axis=1))
    A = V.T @ V
    out.append((0.5 * (A + A.T), V.T @ y))
# cores of 65..128 unknowns (the reference's own benchmark point benchmark_cupy.py:28-36 is r = p = 5: n = 125)
for (rl, p, rr, spread) in [(5, 5, 5, 0.0), (5, 5, 5, 1.5), (4, 6, 4, 0.5), (4, 8, 4, 0.3), (5, 3, 5, 1.0), (5, 5, 4, 0.5)]:
    V = design(rl, p, rr, 2000, spread)
    y = np.log(1 + np.sum(V[:, :3] ** 2, axis=1))
    A = V.T @ V
    out.append((0.5 * (A + A.T), V.T @ y))
for n in (6, 16, 64):
    v = rng.randn(n); out.append((2000.0 * np.outer(v, v), 3.7 * v))
M = rng.randn(18, 20); A = M.T @ M; out.append((A, A @ rng.randn(20)))
for n in (12, 27, 64, 100):
    M = rng.randn(3 * n + 5, n); out.append((M.T @ M, rng.randn(n)))
with open(os.path.join(HERE, "vh_corpus.bin"), "wb") as f:
    f.write(struct.pack("i", len(out)))
    for A, b in out:
        n = A.shape[0]
        x, _, rk, sv = np.linalg.lstsq(A, b, rcond=None)
        f.write(struct.pack("ii", n, int(rk)))
        f.write(np.ascontiguousarray(A, dtype=np.float64).tobytes()); f.write(np.ascontiguousarray(b, dtype=np.float64).tobytes()); f.write(np.ascontiguousarray(x, dtype=np.float64).tobytes())
print("systems", len(out))
