"""TEST INFRASTRUCTURE ONLY (oracle/): stand-in for the un-vendored, un-pinned third-party
package `opt_einsum`, which the reference imports at
/root/reference/src/bsde_solver/core/tensor/tensor_network.py:5 and calls at :82 as
``contract(A, labelsA, B, labelsB, ..., out_labels, optimize='auto')`` (interleaved form,
arbitrary hashable labels).  `opt_einsum` is not installed in this image and there is no
network, so the unmodified reference is run on top of this shim.

Published algorithm restated: a pairwise contraction path chosen greedily by smallest
intermediate, each pair evaluated by ``numpy.einsum`` (BLAS where possible).  Labels are
re-mapped to einsum letters per pair because numpy.einsum only has 52 symbols and a d=100
tensor train carries >100 rank labels.  Two path flavours are offered so the harness can
measure the reference-vs-reference spread: BSDE_SHIM_ORDER=greedy (default) | reverse.
"""
import os
import string
import numpy as np

_LETTERS = string.ascii_letters
_PATH_CACHE = {}


def _pair_einsum(a, la, b, lb, keep):
    """Contract two operands, keeping labels in `keep` (ordered as first seen in la+lb)."""
    out = [l for l in dict.fromkeys(list(la) + list(lb)) if l in keep]
    m = {}
    for l in list(la) + list(lb):
        if l not in m:
            m[l] = _LETTERS[len(m)]
    expr = "".join(m[l] for l in la) + "," + "".join(m[l] for l in lb) + "->" + "".join(m[l] for l in out)
    return np.einsum(expr, a, b, optimize=True), tuple(out)


def _single(a, la, out):
    m = {l: _LETTERS[i] for i, l in enumerate(dict.fromkeys(la))}
    expr = "".join(m[l] for l in la) + "->" + "".join(m[l] for l in out)
    return np.einsum(expr, a)


def _find_path(labels, shapes, out):
    """Greedy smallest-intermediate order over pairs that share a non-batch label."""
    sizes = {}
    for ls, sh in zip(labels, shapes):
        for l, s in zip(ls, sh):
            sizes[l] = max(sizes.get(l, 1), s)
    live = [tuple(ls) for ls in labels]
    ids = list(range(len(live)))
    path = []
    reverse = os.environ.get("BSDE_SHIM_ORDER", "greedy") == "reverse"
    # (label -> number of live operands carrying it: a label survives a pairwise contraction when it is an output label
    #  or some OTHER live operand carries it.  Same choices as the plain O(n^4) search, found ~100x faster, so that the
    #  d = 100 trains of bench.py --impl reference do not spend minutes here.)
    count = {}
    for ls in live:
        for l in set(ls):
            count[l] = count.get(l, 0) + 1
    outset = set(out)
    while len(live) > 1:
        best = None
        n = len(live)
        sets = [set(ls) for ls in live]
        cand = [(i, j) for i in range(n) for j in range(i + 1, n)]
        if reverse:
            cand = cand[::-1]
        for i, j in cand:
            si, sj = sets[i], sets[j]
            shared = (si & sj) - {"batch"}
            res = [l for l in dict.fromkeys(live[i] + live[j])
                   if l in outset or count[l] - (1 if l in si else 0) - (1 if l in sj else 0) > 0]
            size = 1
            for l in res:
                size *= sizes[l]
            key = (0 if shared else 1, size)
            if best is None or key < best[0]:
                best = (key, i, j, tuple(res))
        _, i, j, res = best
        path.append((i, j))
        for l in sets[i]:
            count[l] -= 1
        for l in sets[j]:
            count[l] -= 1
        for l in set(res):
            count[l] = count.get(l, 0) + 1
        live = [live[k] for k in range(n) if k != i and k != j] + [res]
    return path


def contract(*args, optimize="auto", **kwargs):
    args = list(args)
    if len(args) % 2 == 1:
        out = tuple(args.pop())
    else:
        out = None
    ops = [np.asarray(a).view(np.ndarray) for a in args[0::2]]
    labels = [tuple(l) for l in args[1::2]]
    if out is None:
        flat = [l for ls in labels for l in ls]
        out = tuple(l for l in dict.fromkeys(flat) if flat.count(l) == 1)
    key = (tuple(labels), tuple(o.shape for o in ops), out, os.environ.get("BSDE_SHIM_ORDER", "greedy"))
    path = _PATH_CACHE.get(key)
    if path is None:
        path = _find_path(labels, [o.shape for o in ops], out)
        _PATH_CACHE[key] = path
    ops = list(ops)
    labels = list(labels)
    for i, j in path:
        others = set(out)
        for k in range(len(ops)):
            if k != i and k != j:
                others |= set(labels[k])
        r, lr = _pair_einsum(ops[i], labels[i], ops[j], labels[j], others)
        ops = [ops[k] for k in range(len(ops)) if k != i and k != j] + [r]
        labels = [labels[k] for k in range(len(labels)) if k != i and k != j] + [lr]
    return np.ascontiguousarray(_single(ops[0], labels[0], out))


def contract_path(*args, **kwargs):
    raise NotImplementedError("contract_path is not provided by the oracle shim")
