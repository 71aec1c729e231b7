"""Compressed (approximate) contraction on the device -- SURVEY.md 8(f) row f4.

The reference reaches this through ``QuimbBackend(hyperoptimiser=HyperCompressedOptimizer
| ReusableHyperCompressedOptimizer, contraction_params={...})``:
``quimb_tn.contract_compressed(optimize=..., output_inds=..., **contraction_params)``
(optyx/core/backends.py:516-545).  The algorithm (Gray & Chan, "Hyper-optimized
approximate contraction of tensor networks with arbitrary geometry", sec. II): follow a
pairwise contraction order and after every contraction truncate each bond between the
new tensor and a neighbour to at most ``max_bond`` singular values (``cutoff`` relative
to the largest).

Everything numeric runs in our kernels: leaves by ``b2_build_leaves``, pairwise
contractions / Gram matrices / projector applications by ``b2_contract_direct``,
``conj`` by ``b2_conj``, the small bond factorisation by ``b2_bond_compress`` (one-sided
Jacobi, csrc/compress.cu).  The host only decides WHICH pair comes next (sizes, no
values) and reads back one integer per compression (the kept bond dimension, which
fixes the shapes of what follows).  Hyper-indices are reified as COPY tensors first, as
``to_quimb`` does for the reference: compression needs every bond to join exactly two
tensors.  Working precision is complex128; a complex64 backend gets its result cast at
the end.

The Gram route resolves singular values down to ~1e-8 of the largest (it squares the
condition number), so the effective relative cutoff is ``max(cutoff, 3e-8)``; the
reference's own parity statement for this path is ``np.allclose`` (its leaves are even
perturbed by 1e-12 noise first, ``preprocess_quimb_tensors_safe`` utils.py:323-344).
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import runtime as rt
from .core.network import K_DENSE, Leaf, Network
from .planner import _merge_modes

GRAM_FLOOR = 3e-8
MAX_BOND_DEVICE = 256


class _Dev:
    """Device plumbing shared by the steps of one compressed contraction."""

    def __init__(self):
        from .executor import require_cuda
        self.torch = require_cuda()
        self.lib = rt.load()
        self.device = self.torch.device("cuda", self.torch.cuda.current_device())
        self.launches = 0

    @property
    def stream(self):
        return self.torch.cuda.current_stream().cuda_stream

    def empty(self, shape):
        return self.torch.empty(tuple(int(s) for s in shape) or (), dtype=self.torch.complex128,
                                device=self.device)

    def contract(self, a, ia: List[int], b, ib: List[int], out: List[int], dims: Dict[int, int]):
        """``C[out] = sum over the other indices of A[ia] * B[ib]`` (device tensors, any
        strides; every index of ``out`` sits on A or B) through ``b2_contract_direct``."""
        c = self.empty([dims[i] for i in out])

        def strides(t, inds):
            st: Dict[int, int] = {}
            for i, s in zip(inds, t.stride()):
                st[i] = st.get(i, 0) + int(s)           # a repeated index reads the diagonal
            return st
        sa, sb = strides(a, ia), strides(b, ib)
        sc = {i: int(s) for i, s in zip(out, c.stride())} if out else {}
        red = [i for i in dict.fromkeys(list(ia) + list(ib)) if i not in sc]
        outm = _merge_modes([[dims[i], sa.get(i, 0), sb.get(i, 0), sc[i]] for i in out],
                            fastest_last=True)
        redm = _merge_modes(sorted([[dims[i], sa.get(i, 0), sb.get(i, 0), 0] for i in red],
                                   key=lambda m: (-m[1], -m[2])), fastest_last=True)
        if len(redm) > rt.MAX_RED_DIRECT or len(outm) + len(redm) > rt.MAX_MODES:
            raise ValueError("compressed contraction: a pair has too many index groups for the kernel")
        desc = rt.Modes()
        desc.n_out, desc.n_red = len(outm), len(redm)
        for k, m in enumerate(outm + redm):
            desc.dim[k], desc.sa[k], desc.sb[k], desc.sc[k] = m
        rt.check(self.lib.b2_contract_direct(C.byref(desc), a.data_ptr(), b.data_ptr(), c.data_ptr(),
                                             1.0, 0.0, 0, rt.B2_C128, self.stream), "b2_contract_direct")
        self.launches += 1
        return c

    def conj(self, a):
        a = a.contiguous()
        out = self.torch.empty_like(a)
        rt.check(self.lib.b2_conj(a.numel(), a.data_ptr(), out.data_ptr(), rt.B2_C128, self.stream),
                 "b2_conj")
        self.launches += 1
        return out


def _leaf_tensors(dev: _Dev, net: Network):
    """Every leaf array built on the device by the array-build kernel; returns views
    into one arena (complex128)."""
    from .executor import leaf_table
    offs, off = [], 0
    for leaf in net.leaves:
        offs.append(off)
        off += leaf.size
    plan = SimpleNamespace(leaf_offsets=offs, unit_offset=off, arena_elems=off + 1)
    table, pvec = leaf_table(net, plan)
    torch = dev.torch
    desc = torch.from_numpy(table.view(np.uint8).copy()).to(dev.device)
    params = torch.from_numpy(np.ascontiguousarray(pvec)).to(dev.device)
    arena = dev.empty([off + 1])
    rt.check(dev.lib.b2_build_leaves(desc.data_ptr(), len(table), off + 1, params.data_ptr(),
                                     len(pvec), arena.data_ptr(), off + 1, 1, rt.B2_C128, dev.stream),
             "b2_build_leaves")
    dev.launches += 1
    unit = arena[off:off + 1].reshape(())
    return [arena[o:o + leaf.size].reshape(leaf.shape) for o, leaf in zip(offs, net.leaves)], unit


def _compress_bond(dev: _Dev, a, ia, b, ib, dims, fresh: int, max_bond: Optional[int],
                   cutoff: float):
    """Fuse all bonds shared by ``a`` and ``b`` into one index ``fresh`` of at most
    ``max_bond`` values; returns the two new tensors and their index lists."""
    torch = dev.torch
    shared = [i for i in ia if i in ib]
    bond = int(np.prod([dims[i] for i in shared]))
    if bond > MAX_BOND_DEVICE:
        raise ValueError(f"compressed contraction: a bond of dimension {bond} > {MAX_BOND_DEVICE}; "
                         "lower max_bond")
    primed = {i: -1 - k for k, i in enumerate(shared)}       # fresh labels for the bra copy
    pdims = dict(dims)
    for i, p in primed.items():
        pdims[p] = dims[i]
    pr = [primed[i] for i in shared]
    # Gram matrices over everything but the bond: G_A[b', b] = sum conj(A)[.., b'] A[.., b]
    ga = dev.contract(dev.conj(a), [primed.get(i, i) for i in ia], a, ia, pr + shared, pdims)
    # G_B[b, b'] = sum B[b, ..] conj(B)[b', ..]
    gb = dev.contract(b, ib, dev.conj(b), [primed.get(i, i) for i in ib], shared + pr, pdims)
    pa = dev.empty([bond, bond])
    pb = dev.empty([bond, bond])
    k_dev = torch.zeros(1, dtype=torch.int32, device=dev.device)
    sv = torch.empty(bond, dtype=torch.float64, device=dev.device)
    scratch = torch.empty(int(dev.lib.b2_bond_compress_scratch_bytes(bond)), dtype=torch.uint8,
                          device=dev.device)
    rt.check(dev.lib.b2_bond_compress(ga.data_ptr(), gb.data_ptr(), bond,
                                      int(max_bond) if max_bond else 0, max(float(cutoff), GRAM_FLOOR),
                                      pa.data_ptr(), pb.data_ptr(), k_dev.data_ptr(), sv.data_ptr(),
                                      scratch.data_ptr(), dev.stream), "b2_bond_compress")
    dev.launches += 1
    k = int(k_dev.item())                    # the one host read-back: it fixes the shapes below
    dims[fresh] = k
    bshape = [dims[i] for i in shared]
    # P_A[bond, :k] / P_B[:k, bond] as tensors over the bond's own indices (strided views)
    row = [int(np.prod(bshape[j + 1:])) for j in range(len(bshape))]
    pa_t = torch.as_strided(pa, bshape + [k], [r * bond for r in row] + [1])
    pb_t = torch.as_strided(pb, [k] + bshape, [bond] + row)
    rest_a = [i for i in ia if i not in shared]
    rest_b = [i for i in ib if i not in shared]
    a2 = dev.contract(a, ia, pa_t, shared + [fresh], rest_a + [fresh], dims)
    b2 = dev.contract(pb_t, [fresh] + shared, b, ib, rest_b + [fresh], dims)
    dropped = float(0.0)
    return (a2, rest_a + [fresh]), (b2, rest_b + [fresh]), k, dropped


def contract_compressed(net: Network, dtype=np.complex128, max_bond: Optional[int] = None,
                        cutoff: float = 1e-10, return_info: bool = False, **unused):
    """Approximate dense result of ``net`` as a device tensor of shape
    ``net.output_shape``.  ``max_bond`` / ``cutoff`` are quimb's parameter names (the
    reference forwards ``contraction_params`` verbatim, backends.py:539-545); other
    keyword arguments of quimb's ``contract_compressed`` are accepted and ignored."""
    dev = _Dev()
    torch = dev.torch
    dims: Dict[int, int] = dict(net.dims)
    arrays, unit = _leaf_tensors(dev, net)
    tensors: Dict[int, Tuple[object, List[int]]] = {}
    for k, (arr, leaf) in enumerate(zip(arrays, net.leaves)):
        inds = list(leaf.inds)
        while len(set(inds)) != len(inds):            # an index wired twice: the diagonal view
            for i in inds:
                pos = [p for p, j in enumerate(inds) if j == i]
                if len(pos) > 1:
                    arr = torch.diagonal(arr, dim1=pos[0], dim2=pos[1])
                    inds = [j for p, j in enumerate(inds) if p not in pos[:2]] + [i]
                    break
        tensors[k] = (arr, inds)
    out = list(net.output_inds)
    count: Dict[int, int] = {}
    for _, inds in tensors.values():
        for i in inds:
            count[i] = count.get(i, 0) + 1
    for k, (arr, inds) in list(tensors.items()):      # an index nobody else carries is summed out
        lone = [i for i in inds if count[i] == 1 and i not in out]
        if lone:
            keep = [i for i in inds if i not in lone]
            tensors[k] = (dev.contract(arr, inds, unit, [], keep, dims), keep)
    # ---- hyper-indices -> COPY tensors (every index then joins exactly two places)
    nxt_i = (max(dims) + 1) if dims else 0
    nxt_t = (max(tensors) + 1) if tensors else 0
    where: Dict[int, List[int]] = {}
    for k, (_, inds) in tensors.items():
        for i in inds:
            where.setdefault(i, []).append(k)
    for i in sorted(set(where) | set(out)):
        holders = where.get(i, [])
        n_out = out.count(i)
        if len(holders) + n_out <= 2 and holders and n_out <= 1:
            continue
        d = dims[i]
        fresh = []
        for k in holders:
            f, nxt_i = nxt_i, nxt_i + 1
            dims[f] = d
            arr, inds = tensors[k]
            inds = list(inds)
            inds[inds.index(i)] = f
            tensors[k] = (arr, inds)
            fresh.append(f)
        for p, j in enumerate(out):
            if j == i:
                f, nxt_i = nxt_i, nxt_i + 1
                dims[f] = d
                out[p] = f
                fresh.append(f)
        copy = np.zeros((d,) * len(fresh), dtype=np.complex128)
        for v in range(d):
            copy[(v,) * len(fresh)] = 1.0
        tensors[nxt_t] = (torch.from_numpy(copy).to(dev.device), fresh)
        nxt_t += 1
    info = {"max_bond_seen": 0, "compressions": 0, "peak_size": 0, "launches": 0}
    while len(tensors) > 1:
        # greedy: the connected pair with the smallest result; outer products last
        best = None
        keys = sorted(tensors)
        size_of = {k: int(np.prod([dims[i] for i in tensors[k][1]], dtype=np.int64)) for k in keys}
        for x, ka in enumerate(keys):
            ia = tensors[ka][1]
            sa = set(ia)
            for kb in keys[x + 1:]:
                ib = tensors[kb][1]
                if sa.isdisjoint(ib):
                    continue
                size = float(np.prod([float(dims[i]) for i in ia + ib if not (i in sa and i in ib)]))
                score = size - size_of[ka] - size_of[kb]
                if best is None or score < best[0]:
                    best = (score, ka, kb)
        if best is None:
            ka, kb = sorted(tensors, key=lambda k: size_of[k])[:2]
        else:
            _, ka, kb = best
        (a, ia), (b, ib) = tensors.pop(ka), tensors.pop(kb)
        shared = [i for i in ia if i in ib]
        ic = [i for i in ia if i not in shared] + [i for i in ib if i not in shared]
        c = dev.contract(a, ia, b, ib, ic, dims)
        info["peak_size"] = max(info["peak_size"], int(c.numel()))
        for kn in sorted(tensors):                    # compress every bond to a neighbour
            n_arr, n_inds = tensors[kn]
            bonds = [i for i in ic if i in n_inds]
            if not bonds:
                continue
            bond_dim = int(np.prod([dims[i] for i in bonds]))
            info["max_bond_seen"] = max(info["max_bond_seen"], bond_dim)
            if len(bonds) == 1 and (max_bond is None or bond_dim <= max_bond) and cutoff <= 0:
                continue
            (c, ic), (n_arr, n_inds), _, _ = _compress_bond(dev, c, ic, n_arr, n_inds, dims, nxt_i,
                                                           max_bond, cutoff)
            nxt_i += 1
            tensors[kn] = (n_arr, n_inds)
            info["compressions"] += 1
        tensors[nxt_t] = (c, ic)
        nxt_t += 1
    if tensors:
        (result, rinds), = tensors.values()
        if rinds != out:
            result = dev.contract(result, rinds, unit, [], out, dims)     # permute to wire order
    else:
        result = unit
    scal = torch.tensor(complex(net.scalar), dtype=torch.complex128, device=dev.device)
    result = dev.contract(result.reshape([dims[i] for i in out]), out, scal, [], out, dims)
    result = result.reshape(tuple(net.dims[i] for i in net.output_inds))
    if np.dtype(dtype) == np.complex64:
        result = result.to(torch.complex64)
    info["launches"] = dev.launches
    return (result, info) if return_info else result
