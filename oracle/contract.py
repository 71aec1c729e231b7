"""ORACLE (test infrastructure, NOT product code) -- network contraction.

CPU restatement of what the reference does at its two contraction call sites
(optyx/core/backends.py:514 ``quimb_tn ^ ...`` and 539-545
``quimb_tn.contract(optimize=..., output_inds=...)``): contract every tensor
of the network pairwise down to one dense complex128 array whose axes are the
open indices.  quimb/cotengra (third party, un-vendored, no version pinned in
the reference's pyproject.toml:29-39) execute each pairwise step as
transpose -> reshape -> BLAS ``zgemm`` (``numpy.tensordot``/batched
``matmul``); ``pair_contract`` below restates exactly that with numpy.
The pair *order* is free (the result does not depend on it), so the oracle
offers two independent orders -- a size-greedy one and a seeded random one --
which tests require to agree.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import this module.
"""
from __future__ import annotations

import random
from typing import Dict, List, Sequence, Tuple

import numpy as np

from . import leaves as _leaves


def _take_diagonals(arr: np.ndarray, inds: List[int]):
    """A leaf wired twice to the same index means "take the diagonal"."""
    inds = list(inds)
    while len(set(inds)) != len(inds):
        for i in inds:
            pos = [p for p, j in enumerate(inds) if j == i]
            if len(pos) > 1:
                a, b = pos[0], pos[1]
                arr = np.diagonal(arr, axis1=a, axis2=b)   # diag axis goes last
                inds = [j for p, j in enumerate(inds) if p not in (a, b)] + [i]
                break
    return arr, inds


def pair_contract(a: np.ndarray, ia: Sequence[int], b: np.ndarray,
                  ib: Sequence[int], keep: set) -> Tuple[np.ndarray, List[int]]:
    """One pairwise step: sum over indices not in ``keep``; indices in both
    operands *and* in ``keep`` are batch (hyper) indices."""
    a, ia = _take_diagonals(a, list(ia))
    b, ib = _take_diagonals(b, list(ib))
    # indices private to one operand and not kept: sum them out first
    only_a = [i for i in ia if i not in ib and i not in keep]
    if only_a:
        a = a.sum(axis=tuple(ia.index(i) for i in only_a))
        ia = [i for i in ia if i not in only_a]
    only_b = [i for i in ib if i not in ia and i not in keep]
    if only_b:
        b = b.sum(axis=tuple(ib.index(i) for i in only_b))
        ib = [i for i in ib if i not in only_b]
    batch = [i for i in ia if i in ib and i in keep]
    con = [i for i in ia if i in ib and i not in keep]
    m = [i for i in ia if i not in ib]
    n = [i for i in ib if i not in ia]
    dims = {i: d for i, d in zip(ia, a.shape)}
    dims.update({i: d for i, d in zip(ib, b.shape)})

    def prod(xs):
        return int(np.prod([dims[i] for i in xs], dtype=np.int64)) if xs else 1

    a2 = a.transpose([ia.index(i) for i in batch + m + con]).reshape(
        prod(batch), prod(m), prod(con))
    b2 = b.transpose([ib.index(i) for i in batch + con + n]).reshape(
        prod(batch), prod(con), prod(n))
    c2 = np.matmul(a2, b2)          # BLAS zgemm (batched)
    out_inds = batch + m + n
    return c2.reshape([dims[i] for i in out_inds]), out_inds


def _greedy_order(inds_list: List[List[int]], dims: Dict[int, int],
                  output: Sequence[int], rng: random.Random | None):
    """SSA pair order.  ``rng=None``: greedy on size(result) - size(a) - size(b);
    otherwise a seeded random connected pair each step."""
    live = {k: set(v) for k, v in enumerate(inds_list)}
    count: Dict[int, int] = {}           # how many live tensors carry index q
    where: Dict[int, set] = {}
    for k, s in live.items():
        for q in s:
            count[q] = count.get(q, 0) + 1
            where.setdefault(q, set()).add(k)
    out = set(output)
    order = []
    nxt = len(inds_list)

    def result_inds(i, j):
        res = set()
        for q in live[i] | live[j]:
            n_here = (q in live[i]) + (q in live[j])
            if q in out or count[q] > n_here:
                res.add(q)
        return res

    def size(s):
        r = 1.0
        for q in s:
            r *= dims[q]
        return r

    while len(live) > 1:
        cands = set()
        for q, ks in where.items():
            ks = sorted(ks)
            for x in range(len(ks)):
                for y in range(x + 1, len(ks)):
                    cands.add((ks[x], ks[y]))
        if not cands:      # disconnected: outer product of the two smallest
            keys = sorted(live, key=lambda k: size(live[k]))
            cands = {(keys[0], keys[1])}
        cands = sorted(cands)
        if rng is not None:
            i, j = rng.choice(cands)
        else:
            # the usual greedy score: how much the pair shrinks what is alive
            i, j = min(cands, key=lambda p: (size(result_inds(*p)) - size(live[p[0]])
                                             - size(live[p[1]]), p))
        new = result_inds(i, j)
        order.append((i, j, new))
        for k in (i, j):
            for q in live[k]:
                count[q] -= 1
                where[q].discard(k)
            del live[k]
        live[nxt] = new
        for q in new:
            count[q] += 1
            where[q].add(nxt)
        for q in [q for q, ks in where.items() if not ks]:
            del where[q]
        nxt += 1
    return order


def contract_arrays(arrays: List[np.ndarray], inds_list: List[List[int]],
                    dims: Dict[int, int], output_inds: Sequence[int], scalar: complex = 1.0,
                    order: str = "greedy", seed: int = 0) -> np.ndarray:
    """Dense result of a tensor network given as plain arrays + integer index
    labels (an index may sit on any number of tensors and may repeat on one):
    complex128 array of shape ``[dims[i] for i in output_inds]``."""
    arrays, inds_list = list(arrays), [list(ix) for ix in inds_list]
    out_unique = list(dict.fromkeys(output_inds))
    carried = {q for ix in inds_list for q in ix}
    for q in out_unique:
        if q not in carried:             # a bare wire (identity diagram): free end = all ones
            arrays.append(np.ones(dims[q], dtype=complex))
            inds_list.append([q])
    tensors = {k: (np.asarray(arr), list(ix)) for k, (arr, ix) in
               enumerate(zip(arrays, inds_list))}
    if not tensors:
        result, rinds = np.array(1.0 + 0j), []
    else:
        rng = None if order == "greedy" else random.Random(seed)
        ssa = _greedy_order([t[1] for t in tensors.values()], dims,
                            out_unique, rng)
        nxt = len(tensors)
        for i, j, keep in ssa:
            (a, ia), (b, ib) = tensors.pop(i), tensors.pop(j)
            tensors[nxt] = pair_contract(a, ia, b, ib, keep)
            nxt += 1
        (result, rinds), = tensors.values()
        result, rinds = _take_diagonals(result, rinds)
        extra = [i for i in rinds if i not in out_unique]
        if extra:
            result = result.sum(axis=tuple(rinds.index(i) for i in extra))
            rinds = [i for i in rinds if i not in extra]
    result = np.asarray(result, dtype=complex)
    result = result.transpose([rinds.index(i) for i in out_unique]) \
        if rinds else result
    result = result * scalar
    if len(out_unique) != len(output_inds):
        # repeated open index: expand onto the generalised diagonal
        full = np.zeros([dims[i] for i in output_inds], dtype=complex)
        it = np.ndindex(*[dims[i] for i in out_unique])
        for idx in it:
            pos = dict(zip(out_unique, idx))
            full[tuple(pos[i] for i in output_inds)] = result[idx]
        result = full
    return np.require(result, dtype=complex, requirements="C")


def contract_network(net, order: str = "greedy", seed: int = 0,
                     arrays: List[np.ndarray] | None = None) -> np.ndarray:
    """Dense result of ``net`` (an ``optyx_b200.core.network.Network`` or any
    object with ``leaves``, ``dims``, ``output_inds``, ``scalar``): complex128
    array of shape ``[dims[i] for i in output_inds]``."""
    if arrays is None:
        arrays = _leaves.build_all(net.leaves)
    return contract_arrays(arrays, [list(l.inds) for l in net.leaves], net.dims,
                           net.output_inds, net.scalar, order, seed)


def contract_along_path(arrays: List[np.ndarray], inds_list: List[List[int]],
                        dims: Dict[int, int], output_inds: Sequence[int],
                        ssa_path: Sequence[Tuple[int, int]], fixed: Dict[int, int] | None = None,
                        scalar: complex = 1.0) -> np.ndarray:
    """The same pairwise arithmetic along a GIVEN SSA pair order, with the indices in
    ``fixed`` pinned to one value each (one SLICE of a sliced contraction, as cotengra's
    ``contract_slice`` evaluates it; the caller sums the slices).  Used to time the CPU
    path on a bounded sample of a sliced workload and to check sliced execution."""
    fixed = fixed or {}
    tensors = {}
    for k, (arr, ix) in enumerate(zip(arrays, inds_list)):
        arr, ix = np.asarray(arr), list(ix)
        for q, v in fixed.items():
            while q in ix:
                p = ix.index(q)
                arr = np.take(arr, v, axis=p)
                ix.pop(p)
        tensors[k] = (arr, ix)
    out_unique = list(dict.fromkeys(output_inds))
    count: Dict[int, int] = {}
    for _, ix in tensors.values():
        for q in set(ix):
            count[q] = count.get(q, 0) + 1
    nxt = len(tensors)
    for i, j in ssa_path:
        (a, ia), (b, ib) = tensors.pop(i), tensors.pop(j)
        keep = {q for q in set(ia) | set(ib)
                if q in out_unique or count[q] > (q in ia) + (q in ib)}
        for q in set(ia):
            count[q] -= 1
        for q in set(ib):
            count[q] -= 1
        c, ic = pair_contract(a, ia, b, ib, keep)
        for q in set(ic):
            count[q] = count.get(q, 0) + 1
        tensors[nxt] = (c, ic)
        nxt += 1
    (result, rinds), = tensors.values()
    result, rinds = _take_diagonals(result, rinds)
    extra = [i for i in rinds if i not in out_unique]
    if extra:
        result = result.sum(axis=tuple(rinds.index(i) for i in extra))
        rinds = [i for i in rinds if i not in extra]
    result = np.asarray(result, dtype=complex)
    if rinds:
        result = result.transpose([rinds.index(i) for i in out_unique])
    return result * scalar
