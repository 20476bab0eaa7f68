"""Shared helpers of the parity tests."""
import numpy as np


def cores_from(npz, prefix):
    out, i = [], 0
    while f"{prefix}{i}" in npz:
        out.append(np.array(npz[f"{prefix}{i}"]))
        i += 1
    return out


def ranks_of(cores):
    return np.array([c.shape[0] for c in cores] + [cores[-1].shape[2]], dtype=np.int32)


def flat(cores):
    return np.concatenate([np.ascontiguousarray(c, dtype=np.float64).ravel() for c in cores])


def unflat(flat_arr, ranks, p):
    out, off = [], 0
    for i in range(len(ranks) - 1):
        n = int(ranks[i]) * p * int(ranks[i + 1])
        out.append(np.array(flat_arr[off:off + n]).reshape(int(ranks[i]), p, int(ranks[i + 1])))
        off += n
    return out


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
