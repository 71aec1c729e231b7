"""Host-side contraction planner: pair order, index slicing, operand layouts and
the per-step mode descriptors the CUDA kernels consume.

The reference delegates all of this to quimb/cotengra (``quimb_tn ^ ...`` and
``contract(optimize=hyperoptimiser)``, optyx/core/backends.py:513-545), neither
of which is available here; this module is a from-scratch planner designed
around the B200 kernels rather than around numpy:

* hyper-indices are first class (a spider is a shared index, never a COPY
  tensor), so a pair can have *batch* modes besides M/N/K;
* no tensor is ever transposed in HBM: every step reads its operands through
  per-mode strides (the permutation is fused into the gather) and the planner
  chooses each intermediate's layout so that the big operand streams;
* slicing fixes index values by pointer offsets into the leaf arena, so slices
  cost no copies and slice-invariant subtrees are contracted once.
"""
from __future__ import annotations

import heapq
import math
import os
import random
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .core.network import Network
from .runtime import MAX_MODES, MAX_RED_DIRECT, GemmModes, Modes


# --------------------------------------------------------------------------
# pair order
# --------------------------------------------------------------------------
def _prod(dims: Dict[int, int], inds) -> float:
    r = 1.0
    for i in inds:
        r *= dims[i]
    return r


def greedy_path(inputs: Sequence[frozenset], dims: Dict[int, int], output: frozenset,
                rng: Optional[random.Random] = None, temperature: float = 0.0,
                costmod: float = 1.0) -> List[Tuple[int, int]]:
    """SSA path (leaf ids 0..n-1, step s creates id n+s) by greedy pairing on
    ``size(out) - costmod * (size(a) + size(b))`` with optional Gumbel noise."""
    n = len(inputs)
    tens: Dict[int, frozenset] = {k: frozenset(v) for k, v in enumerate(inputs)}
    count: Dict[int, int] = {}
    where: Dict[int, set] = {}
    for k, s in tens.items():
        for q in s:
            count[q] = count.get(q, 0) + 1
            where.setdefault(q, set()).add(k)
    for q in output:
        count[q] = count.get(q, 0) + 1
    size = {k: _prod(dims, s) for k, s in tens.items()}

    def out_inds(a: int, b: int) -> frozenset:
        sa, sb = tens[a], tens[b]
        res = []
        for q in sa | sb:
            here = (q in sa) + (q in sb)
            if count[q] > here:
                res.append(q)
        return frozenset(res)

    heap: List[tuple] = []
    tick = 0

    def push(a: int, b: int) -> None:
        nonlocal tick
        so = _prod(dims, out_inds(a, b))
        sc = so - costmod * (size[a] + size[b])
        sc = math.copysign(math.log2(1.0 + abs(sc)), sc)
        if rng is not None and temperature > 0.0:
            u = rng.random()
            sc -= temperature * -math.log(-math.log(max(u, 1e-300)))
        tick += 1
        heapq.heappush(heap, (sc, tick, a, b))

    def neighbours(a: int) -> set:
        res = set()
        for q in tens[a]:
            ks = where[q]
            if len(ks) > 24:           # huge hyper-edge: a few smallest partners do
                ks = sorted(ks, key=lambda k: size[k])[:24]
            res.update(ks)
        res.discard(a)
        return res

    for a in range(n):
        for b in neighbours(a):
            if a < b:
                push(a, b)

    path: List[Tuple[int, int]] = []
    nxt = n
    while len(tens) > 1:
        pair = None
        while heap:
            _, _, a, b = heapq.heappop(heap)
            if a in tens and b in tens:
                pair = (a, b)
                break
        if pair is None:                 # disconnected components: outer products
            ks = sorted(tens, key=lambda k: size[k])
            pair = (ks[0], ks[1])
        a, b = pair
        new = out_inds(a, b)
        for k in (a, b):
            for q in tens[k]:
                count[q] -= 1
                where[q].discard(k)
            del tens[k]
        tens[nxt] = new
        size[nxt] = _prod(dims, new)
        for q in new:
            count[q] += 1
            where.setdefault(q, set()).add(nxt)
        path.append((a, b))
        for c in neighbours(nxt):
            push(nxt, c)
        nxt += 1
    return path


# cost model of the path search: a step costs max(bytes / HBM, flops / FP64 pipe)
# (complex128 on B200: 16 B per element, 6.5 TB/s, 37 TFLOP/s)
_COST_BW, _COST_PIPE, _COST_ESIZE = 6.5e12, 37e12, 16.0


@dataclass
class PathInfo:
    path: List[Tuple[int, int]]
    sliced: Tuple[int, ...]
    macs: float                 # complex multiply-adds per slice
    peak: float                 # largest intermediate (elements) per slice
    nslices: int
    step_inds: List[frozenset] = field(default_factory=list)   # result inds per step
    time_est: float = 0.0       # roofline estimate of the whole (sliced) contraction, s


def analyse_path(inputs: Sequence[frozenset], dims: Dict[int, int], output: frozenset,
                 path: Sequence[Tuple[int, int]], sliced: Sequence[int] = ()) -> PathInfo:
    sl = set(sliced)
    tens = {k: frozenset(v) for k, v in enumerate(inputs)}
    count: Dict[int, int] = {}
    for s in tens.values():
        for q in s:
            count[q] = count.get(q, 0) + 1
    for q in output:
        count[q] = count.get(q, 0) + 1
    macs, peak = 0.0, 0.0
    step_inds = []
    nxt = len(inputs)
    nslices = int(_prod(dims, sl)) if sl else 1
    dep = {k: bool(s & sl) for k, s in tens.items()}      # depends on a sliced index
    t_inv = t_dep = 0.0
    for a, b in path:
        sa, sb = tens.pop(a), tens.pop(b)
        union = sa | sb
        new = frozenset(q for q in union if count[q] > (q in sa) + (q in sb))
        for q in sa:
            count[q] -= 1
        for q in sb:
            count[q] -= 1
        for q in new:
            count[q] += 1
        m = _prod(dims, [q for q in union if q not in sl])
        size_c = _prod(dims, [q for q in new if q not in sl])
        nbytes = _COST_ESIZE * (_prod(dims, [q for q in sa if q not in sl]) +
                                _prod(dims, [q for q in sb if q not in sl]) + size_c)
        t = max(nbytes / _COST_BW, 8.0 * m / _COST_PIPE) + 2e-6      # + launch floor
        d = dep.pop(a) | dep.pop(b)
        dep[nxt] = d
        if d:
            t_dep += t
        else:
            t_inv += t
        macs += m
        peak = max(peak, size_c)
        tens[nxt] = new
        step_inds.append(new)
        nxt += 1
    return PathInfo(list(path), tuple(sliced), macs, peak, nslices, step_inds,
                    t_inv + nslices * t_dep)


NATIVE_MIN_TENSORS = 64        # from here on the search runs in C++ (csrc/planner_native.cu)
NATIVE_TRIALS_PER_TRIAL = 16   # ... and affords this many more randomised trials
NATIVE_SLICE_CANDIDATES = int(os.environ.get("B2_SLICE_CANDS", "48"))   # trees taken through slicing
SLICE_RECONF_ROUNDS = int(os.environ.get("B2_SLICE_ROUNDS", "4"))
RECONF_LEAVES = 8              # frontier size of the subtree reconfiguration


def find_path(inputs, dims, output, trials: int = 8, seed: int = 0,
              minimize: str = "flops", target_size: Optional[float] = None,
              slice_candidates: int = 4, min_slices: int = 1) -> PathInfo:
    """Best of a deterministic greedy plus ``trials-1`` noisy ones.  With a
    ``target_size`` the candidates are compared AFTER slicing, by the roofline
    estimate of the whole sliced contraction (a tree that is cheapest unsliced can
    be far from the cheapest once its widest tensors must be cut): the
    ``slice_candidates`` cheapest trees are sliced and the best sliced one wins.
    ``min_slices``: keep slicing (cheapest index first) until there are at least that
    many slices -- the units that shard over GPUs.  Networks of ``NATIVE_MIN_TENSORS``
    tensors or more are searched by the C++ implementation with 16x the trials."""
    need_slicing = target_size is not None or min_slices > 1
    if len(inputs) >= NATIVE_MIN_TENSORS:
        return _find_path_native(inputs, dims, output, max(1, trials) * NATIVE_TRIALS_PER_TRIAL,
                                 seed, minimize, target_size, max(slice_candidates, NATIVE_SLICE_CANDIDATES), min_slices)
    cands = []
    rng = random.Random(seed)
    for t in range(max(1, trials)):
        if t == 0:
            p = greedy_path(inputs, dims, output)
        else:
            p = greedy_path(inputs, dims, output, rng=rng,
                            temperature=0.3 + 0.7 * rng.random(),
                            costmod=rng.choice([0.5, 1.0, 1.0, 2.0]))
        info = analyse_path(inputs, dims, output, p)
        key = (info.macs, info.peak) if minimize == "flops" else (info.peak, info.macs)
        cands.append((key, t, info))
    cands.sort(key=lambda c: (c[0], c[1]))
    if not need_slicing or (min_slices <= 1 and cands[0][2].peak <= target_size):
        return cands[0][2]
    best = None
    for _, _, info in cands[:max(1, slice_candidates)]:
        sl = slice_path(inputs, dims, output, info, target_size, min_slices=min_slices)
        if best is None or sl.time_est < best.time_est:
            best = sl
    return best


def _find_path_native(inputs, dims, output, trials, seed, minimize, target_size,
                      slice_candidates, min_slices) -> PathInfo:
    nn = NativeNetwork(inputs, dims, output)
    rng = random.Random(seed)
    # the trial parameters are drawn in order (the result does not depend on the thread
    # count); the C++ searches run on a thread pool (ctypes releases the GIL)
    knobs = [(0.0, 1.0)] + [(0.3 + 0.7 * rng.random(), rng.choice([0.5, 1.0, 1.0, 2.0]))
                            for _ in range(1, trials)]

    def one_trial(t):
        if t == 0:
            p = nn.greedy(seed=seed)
        else:
            p = nn.greedy(seed=seed * 7919 + t, temperature=knobs[t][0], costmod=knobs[t][1])
        macs, peak, _, _ = nn.cost(p)
        return (macs, peak, t, p)

    workers = max(1, min(16, os.cpu_count() or 1, trials))
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=workers) as pool:
        cands = list(pool.map(one_trial, range(trials)))
    need_slicing = target_size is not None or min_slices > 1

    def info_of(p, sliced=()):
        return analyse_path(inputs, dims, output, [tuple(x) for x in p.tolist()], sliced)
    # subtree reconfiguration (optimal pair order inside every subtree of <= 8 frontier
    # nodes) on the most promising trees: 3x fewer multiply-adds on cfg4
    tgt = target_size if target_size is not None else float("inf")
    cands.sort(key=lambda c: (c[0] * max(1.0, c[1] / tgt), c[2]))
    def refine(c):
        p = nn.reconfigure(c[3], (), RECONF_LEAVES)
        macs, peak, _, _ = nn.cost(p)
        return (macs, peak, c[2], p)

    with ThreadPoolExecutor(max_workers=workers) as pool:
        refined = list(pool.map(refine, cands[:max(16, 2 * slice_candidates)]))
    if not need_slicing:
        refined.sort(key=lambda c: (c[0], c[1], c[2]) if minimize == "flops" else (c[1], c[0], c[2]))
        return info_of(refined[0][3])
    # rank by work x how far the widest tensor overshoots (each factor of two over the
    # target costs at least one sliced index), slice the best few, re-optimise the tree
    # for the sliced network, compare by the roofline time estimate
    refined.sort(key=lambda c: (c[0] * max(1.0, c[1] / tgt), c[2]))
    def slice_one(c):
        p = c[3]
        info = info_of(p)
        if info.peak > tgt or min_slices > 1:
            info = slice_path(inputs, dims, output, info, target_size, min_slices=min_slices)
            # alternate: re-order the tree for the sliced network, re-choose the slices for
            # the re-ordered tree, ... while the roofline estimate keeps improving
            for _ in range(SLICE_RECONF_ROUNDS):
                p2 = nn.reconfigure(p, info.sliced, RECONF_LEAVES)
                options = [info_of(p2, info.sliced),
                           slice_path(inputs, dims, output, info_of(p2), target_size,
                                      min_slices=min_slices)]
                options = [o for o in options if o.peak <= tgt and o.nslices >= min_slices]
                if not options:
                    break
                again = min(options, key=lambda o: o.time_est)
                if again.time_est >= 0.99 * info.time_est:
                    break
                info, p = again, p2
        return info

    with ThreadPoolExecutor(max_workers=workers) as pool:
        infos = list(pool.map(slice_one, refined[:max(1, slice_candidates)]))
    best = None
    for info in infos:                    # first of equals wins: independent of the thread count
        if best is None or info.time_est < best.time_est:
            best = info
    return best


def slice_path(inputs, dims, output, info: PathInfo, target_size: Optional[float],
               max_slices: int = 1 << 24, min_slices: int = 1) -> PathInfo:
    """Greedy index slicing until the largest intermediate fits ``target_size``
    elements.  Works on step x index incidence matrices (numpy): each round
    scores every candidate index by (bits of oversize removed over all oversized
    intermediates) per (bit of total-MAC overhead) and fixes the best one."""
    path = info.path
    if not path:
        return info
    ids = sorted({q for s in inputs for q in s})
    col = {q: k for k, q in enumerate(ids)}
    logd = np.array([math.log2(dims[q]) for q in ids])
    n_steps = len(path)
    U = np.zeros((n_steps, len(ids)), dtype=bool)      # union of the pair's indices
    R = np.zeros((n_steps, len(ids)), dtype=bool)      # result indices
    tens = {k: frozenset(v) for k, v in enumerate(inputs)}
    nxt = len(inputs)
    for k, ((a, b), res) in enumerate(zip(path, info.step_inds)):
        for q in tens.pop(a) | tens.pop(b):
            U[k, col[q]] = True
        for q in res:
            R[k, col[q]] = True
        tens[nxt] = res
        nxt += 1
    Uf, Rf = U.astype(np.float64), R.astype(np.float64)
    allowed = np.ones(len(ids), dtype=bool)
    for q in output:
        if q in col:
            allowed[col[q]] = False
    allowed &= logd > 0
    cur_u, cur_r = Uf @ logd, Rf @ logd
    log_target = math.log2(target_size) if target_size is not None else float("inf")
    log_min = math.log2(max(1, min_slices))
    sliced: List[int] = []
    log_slices = 0.0
    while (cur_r.max() > log_target + 1e-9 or log_slices < log_min - 1e-9) \
            and log_slices < math.log2(max_slices):
        over = cur_r > log_target + 1e-9
        if over.any():
            cand = np.flatnonzero(allowed & R[over].any(axis=0))
        else:
            # everything fits already: more slices are wanted as units of parallel
            # work -- any index will do, the cheapest (least extra work) wins
            cand = np.flatnonzero(allowed & U.any(axis=0))
        if cand.size == 0:
            break
        excess_now = np.maximum(cur_r - log_target, 0.0).sum()
        macs_now = np.logaddexp2.reduce(cur_u) + log_slices
        # candidate scores, vectorised over candidates
        new_r = cur_r[:, None] - Rf[:, cand] * logd[cand][None, :]
        excess = np.maximum(new_r - log_target, 0.0).sum(axis=0)
        new_u = cur_u[:, None] - Uf[:, cand] * logd[cand][None, :]
        macs = np.logaddexp2.reduce(new_u, axis=0) + log_slices + logd[cand]
        gain = excess_now - excess
        overhead = np.maximum(macs - macs_now, 0.0)
        if over.any():
            score = gain / (overhead + 0.05)
            j = int(np.argmax(score))
            if gain[j] <= 0:
                break
        else:
            j = int(np.argmin(overhead))
        c = int(cand[j])
        sliced.append(ids[c])
        allowed[c] = False
        cur_r = cur_r - Rf[:, c] * logd[c]
        cur_u = cur_u - Uf[:, c] * logd[c]
        log_slices += logd[c]
    return analyse_path(inputs, dims, output, path, sliced)


# --------------------------------------------------------------------------
# layouts and step compilation
# --------------------------------------------------------------------------
ARENA, WS, OUT = 0, 1, 2
PARALLEL_WS_BYTES = 256 << 20


@dataclass
class Layout:
    """A tensor in device memory: unique indices with element strides."""
    inds: Tuple[int, ...]
    strides: Tuple[int, ...]
    space: int                   # ARENA / WS / OUT
    offset: int                  # element offset inside that space (per eval)
    size: int                    # elements per eval (for the batch stride)

    def stride_of(self, q: int) -> int:
        return self.strides[self.inds.index(q)] if q in self.inds else 0


@dataclass
class Step:
    kernel: str                  # "direct" | "gemm"
    desc: object                 # runtime.Modes or runtime.GemmModes
    a: Layout
    b: Layout
    c: Layout
    a_slice: List[Tuple[int, int]]   # (position in plan.sliced, stride) terms
    b_slice: List[Tuple[int, int]]
    per_slice: bool              # False: slice-invariant, run once
    final: bool
    macs: float                  # complex MACs (incl. eval batch)
    bytes: float                 # algorithmic bytes: s * (|A| + |B| + |C|)
    mnk: Tuple[int, int, int, int]   # (batch, M, N, K) extents
    grouped: bool = False        # runs inside the small-step program (plan.small)
    small_modes: Optional[tuple] = None   # per-eval (out modes, red modes) [dim, sa, sb, sc]


@dataclass
class Plan:
    steps: List[Step]
    sliced: Tuple[int, ...]
    sliced_dims: Tuple[int, ...]
    nslices: int
    leaf_offsets: List[int]
    arena_elems: int             # per eval, includes the unit element
    unit_offset: int
    ws_elems: int                # per eval
    out_shape: Tuple[int, ...]
    out_elems: int
    needs_zero_out: bool
    batch: int
    dtype: np.dtype
    info: PathInfo
    total_macs: float = 0.0
    total_bytes: float = 0.0
    parallel_safe: bool = False   # no workspace recycling: branches may overlap
    small: Optional[object] = None        # SmallProgram: every early tiny step, ONE launch


class _Allocator:
    """First-fit free-list allocator over element offsets (plan time only)."""

    def __init__(self):
        self.free: List[Tuple[int, int]] = []     # (offset, size)
        self.top = 0

    def alloc(self, size: int) -> int:
        for k, (off, sz) in enumerate(self.free):
            if sz >= size:
                if sz == size:
                    self.free.pop(k)
                else:
                    self.free[k] = (off + size, sz - size)
                return off
        off = self.top
        self.top += size
        return off

    def release(self, off: int, size: int) -> None:
        self.free.append((off, size))
        self.free.sort()
        merged: List[Tuple[int, int]] = []
        for o, s in self.free:
            if merged and merged[-1][0] + merged[-1][1] == o:
                merged[-1] = (merged[-1][0], merged[-1][1] + s)
            else:
                merged.append((o, s))
        self.free = merged


def _merge_modes(modes: List[List[int]], fastest_last: bool) -> List[List[int]]:
    """Fuse neighbouring modes [dim, sa, sb, sc] whose strides nest in all three
    tensors.  ``fastest_last``: list is slow -> fast (C order)."""
    modes = [m for m in modes if m[0] > 1]
    if not modes:
        return []
    if not fastest_last:
        modes = modes[::-1]
    out = [list(modes[0])]
    for m in modes[1:]:
        slow = out[-1]
        if all(slow[k] == m[k] * m[0] for k in (1, 2, 3)):
            out[-1] = [slow[0] * m[0], m[1], m[2], m[3]]
        else:
            out.append(list(m))
    if not fastest_last:
        out = out[::-1]
    return out


def compile_plan(net: Network, info: PathInfo, dtype=np.complex128, batch: int = 1,
                 gemm_threshold: Tuple = (32, 16, 8, 4.0)) -> Plan:
    """Turn a pair order into executable steps (layouts, descriptors, memory)."""
    dtype = np.dtype(dtype)
    esize = dtype.itemsize
    dims = dict(net.dims)
    sliced = tuple(info.sliced)
    sl = set(sliced)
    out_unique = tuple(dict.fromkeys(net.output_inds))
    assert not (sl & set(out_unique)), "open indices are never sliced"

    # ---- leaf layouts inside the arena (C order, repeated inds -> summed strides)
    leaf_offsets, layouts = [], {}
    off = 0
    for k, leaf in enumerate(net.leaves):
        leaf_offsets.append(off)
        st = {}
        s = 1
        for d, q in zip(reversed(leaf.shape), reversed(leaf.inds)):
            st[q] = st.get(q, 0) + s
            s *= d
        inds = tuple(dict.fromkeys(leaf.inds))
        layouts[k] = Layout(inds, tuple(st[q] for q in inds), ARENA, off, leaf.size)
        off += leaf.size
    unit_offset = off
    arena_elems = off + 1
    unit = Layout((), (), ARENA, unit_offset, 1)

    # ---- which SSA ids depend on a sliced index
    n = len(net.leaves)
    dep = {k: bool(set(net.leaves[k].inds) & sl) for k in range(n)}

    path = list(info.path)
    ids_live = list(range(n))
    if n == 0:
        path_ops = [(None, None)]
    elif n == 1:
        path_ops = [(0, None)]
    else:
        path_ops = list(path)

    count: Dict[int, int] = {}
    for leaf in net.leaves:
        for q in set(leaf.inds):
            count[q] = count.get(q, 0) + 1
    for q in out_unique:
        count[q] = count.get(q, 0) + 1

    alloc = _Allocator()
    steps: List[Step] = []
    nxt = n
    total_macs = total_bytes = 0.0

    # final output strides (repeated open index -> summed strides)
    out_shape = tuple(dims[q] for q in net.output_inds)
    out_st: Dict[int, int] = {}
    s = 1
    for d, q in zip(reversed(out_shape), reversed(net.output_inds)):
        out_st[q] = out_st.get(q, 0) + s
        s *= d
    out_elems = int(s)
    needs_zero = len(out_unique) != len(net.output_inds)

    for si, (ia, ib) in enumerate(path_ops):
        final = si == len(path_ops) - 1
        la = layouts.pop(ia) if ia is not None else unit
        lb = layouts.pop(ib) if ib is not None else unit
        da = dep.get(ia, False) if ia is not None else False
        db = dep.get(ib, False) if ib is not None else False
        if lb.size > la.size:              # A := the larger operand (it streams)
            la, lb, da, db, ia, ib = lb, la, db, da, ib, ia
        sa_set, sb_set = set(la.inds), set(lb.inds)
        keep = []
        for q in la.inds + tuple(x for x in lb.inds if x not in sa_set):
            here = (q in sa_set) + (q in sb_set)
            if count[q] > here and q not in sl:
                keep.append(q)
        for q in sa_set:
            count[q] -= 1
        for q in sb_set:
            count[q] -= 1
        for q in keep:
            count[q] += 1
        keep_set = set(keep)

        def by_stride(lay: Layout, qs):
            return sorted(qs, key=lambda q: -lay.stride_of(q))

        batch_m = by_stride(la, [q for q in keep if q in sa_set and q in sb_set])
        m_m = by_stride(la, [q for q in keep if q in sa_set and q not in sb_set])
        n_m = by_stride(lb, [q for q in keep if q in sb_set and q not in sa_set])
        red = [q for q in la.inds if q not in keep_set and q not in sl] + \
              [q for q in lb.inds if q not in keep_set and q not in sl and q not in sa_set]

        # ---- kernel class (decides the output layout): tensor-core GEMM only for
        # pairs that really are dense (see below), everything else is "thin"
        Mx0, Nx0, Kx0 = _prod(dims, m_m), _prod(dims, n_m), _prod(dims, red)
        Bx0 = _prod(dims, batch_m) * batch
        size_abc = (_prod(dims, [q for q in la.inds if q not in sl]) +
                    _prod(dims, [q for q in lb.inds if q not in sl]) + _prod(dims, keep)) * batch
        ai0 = 8.0 * Bx0 * Mx0 * Nx0 * Kx0 / max(esize * size_abc, 1.0)
        min_ai0 = gemm_threshold[3] if len(gemm_threshold) > 3 else 0.0
        thin = not (Mx0 >= gemm_threshold[0] and Nx0 >= gemm_threshold[1]
                    and Kx0 >= gemm_threshold[2] and ai0 >= min_ai0)
        # ---- output layout
        if final:
            assert set(keep) == set(out_unique), (keep, out_unique)
            c_inds = out_unique
            c_strides = tuple(out_st[q] for q in c_inds)
            lc = Layout(c_inds, c_strides, OUT, 0, out_elems)
        else:
            c_inds = tuple(batch_m + m_m + n_m)
            if thin:
                # bandwidth-bound pair (big A x tiny B): C keeps A's own mode order
                # (minus the reduced modes) so A streams in and C streams out in
                # the same sequence, and the few new modes B brings are the
                # SLOWEST ones -- one contiguous output stream per column instead
                # of 16 B stores strided by N
                c_inds = tuple(n_m + by_stride(la, batch_m + m_m))
            c_strides, s = [], 1
            for q in reversed(c_inds):
                c_strides.append(s)
                s *= dims[q]
            c_strides = tuple(reversed(c_strides))
            size_c = int(s)
            lc = Layout(c_inds, c_strides, WS, alloc.alloc(size_c * batch), size_c)

        ebatch = [[batch, la.size if la.space != ARENA else arena_elems,
                   lb.size if lb.space != ARENA else arena_elems, lc.size]] if batch > 1 else []
        # ---- descriptors
        Bx = int(_prod(dims, batch_m)) * batch
        Mx, Nx, Kx = int(_prod(dims, m_m)), int(_prod(dims, n_m)), int(_prod(dims, red))
        macs = float(Bx) * Mx * Nx * Kx
        sizeA = _prod(dims, [q for q in la.inds if q not in sl])
        sizeB = _prod(dims, [q for q in lb.inds if q not in sl])
        sizeC = _prod(dims, keep)
        nbytes = esize * batch * (sizeA + sizeB + sizeC)

        def mode(q):
            return [dims[q], la.stride_of(q), lb.stride_of(q), lc.stride_of(q)]

        # tensor-core path only for pairs that really are dense GEMMs: big enough
        # in every extent AND above the FP64 ridge (~5.7 flop/B on B200: 37 TF/s
        # over 6.5 TB/s); everything else is bandwidth work for the direct kernel
        intensity = 8.0 * macs / max(nbytes, 1.0)
        min_ai = gemm_threshold[3] if len(gemm_threshold) > 3 else 0.0
        # (complex128: FP64 DMMA; complex64: tcgen05 kind::tf32 split precision)
        use_gemm = (Mx >= gemm_threshold[0]
                    and Nx >= gemm_threshold[1] and Kx >= gemm_threshold[2]
                    and intensity >= min_ai)
        # small output over a long reduction (two big intermediates sharing almost all of
        # their modes): the thin split-K kernel behind b2_contract_gemm reads both operands at
        # HBM speed; the direct kernels serialise the reduction (cfg3: (16, 16, K = 16384) took
        # 237 us on the reduce kernel)
        if (not use_gemm and not final and 4 <= Mx <= 32 and 4 <= Nx <= 32 and Kx >= 256
                and np.dtype(dtype) == np.complex128):
            use_gemm = True
        desc = None
        if not use_gemm:
            # out modes in C *memory* order so that stores are coalesced
            c_order = sorted(c_inds, key=lambda q: -lc.stride_of(q))
            outm = _merge_modes(ebatch + [mode(q) for q in c_order], fastest_last=True)
            redm = sorted([mode(q) for q in red], key=lambda m: (-m[1], -m[2]))
            redm = _merge_modes(redm, fastest_last=True)
            if len(redm) > MAX_RED_DIRECT:
                use_gemm = True
            if not use_gemm:
                if len(outm) + len(redm) > MAX_MODES:
                    raise ValueError("contraction step has too many modes for the kernel")
                desc = Modes()
                desc.n_out, desc.n_red = len(outm), len(redm)
                for k, m in enumerate(outm + redm):
                    desc.dim[k], desc.sa[k], desc.sb[k], desc.sc[k] = m
        if use_gemm:
            gb = _merge_modes(ebatch[:] + [mode(q) for q in batch_m], fastest_last=True)[::-1]
            gm = _merge_modes(sorted([mode(q) for q in m_m], key=lambda m: m[3]), False)
            gn = _merge_modes(sorted([mode(q) for q in n_m], key=lambda m: m[3]), False)
            gk = _merge_modes(sorted([mode(q) for q in red],
                                     key=lambda m: (m[1] if m[1] else 1 << 62, m[2])), False)
            allm = gb + gm + gn + gk
            if len(allm) > MAX_MODES:
                raise ValueError("contraction step has too many modes for the kernel")
            desc = GemmModes()
            desc.n_batch, desc.n_m, desc.n_n, desc.n_k = len(gb), len(gm), len(gn), len(gk)
            for k, m in enumerate(allm):
                desc.dim[k], desc.sa[k], desc.sb[k], desc.sc[k] = m

        small_modes = None
        if not final and max(la.size, lb.size, lc.size) <= SMALL_ELEMS:
            small_modes = ([mode(q) for q in sorted(c_inds, key=lambda q: -lc.stride_of(q))],
                           [mode(q) for q in red])
        a_slice = [(sliced.index(q), la.stride_of(q)) for q in la.inds if q in sl]
        b_slice = [(sliced.index(q), lb.stride_of(q)) for q in lb.inds if q in sl]
        per_slice = bool(da or db)
        steps.append(Step("gemm" if use_gemm else "direct", desc, la, lb, lc,
                          a_slice, b_slice, per_slice, final, macs, nbytes,
                          (Bx, Mx, Nx, Kx), small_modes=small_modes))
        mult = info.nslices if per_slice else 1
        total_macs += macs * mult
        total_bytes += nbytes * mult
        # release operands living in the workspace
        for lay in (la, lb):
            if lay.space == WS:
                alloc.release(lay.offset, lay.size * batch)
        layouts[nxt] = lc
        dep[nxt] = per_slice
        nxt += 1

    # a slice-invariant buffer must survive all slices: the simple allocator
    # above may have recycled it for a per-slice tensor.  Re-run allocation in
    # two arenas when slicing is on.
    parallel_safe = False
    if sliced:
        _reallocate_for_slices(steps, batch)
    else:
        # small plans: give every intermediate a private buffer so independent
        # subtrees may run concurrently (parallel branches of one CUDA graph)
        total = sum(st.c.size * batch for st in steps if st.c.space == WS)
        if total * esize <= PARALLEL_WS_BYTES:
            off = 0
            for st in steps:
                if st.c.space == WS:
                    st.c.offset = off
                    off += st.c.size * batch
            parallel_safe = True
    ws_elems = max([st.c.offset + st.c.size * batch for st in steps if st.c.space == WS] or [0])

    # the small-step program runs before every other launch: what it hands over
    # (task exports) lives in private, never recycled workspace
    small, ws_elems = _build_small_program(steps, batch, arena_elems, int(ws_elems))
    if small is not None:
        for k in small.step_ids:
            steps[k].grouped = True

    plan = Plan(steps, sliced, tuple(dims[q] for q in sliced), info.nslices, leaf_offsets,
                arena_elems, unit_offset, int(ws_elems), out_shape, out_elems,
                needs_zero or bool(sliced), batch, dtype, info, total_macs, total_bytes)
    plan.parallel_safe = parallel_safe
    plan.small = small
    return plan


# --------------------------------------------------------------------------
# the small-step program (csrc/contract_small.cu)
# --------------------------------------------------------------------------
SMALL_ELEMS = 4096                  # largest operand / result of a step the interpreter takes
SMALL_WORK = 1 << 18                # ... and its multiply-adds
SMALL_SCRATCH_ELEMS = 11264         # scratch budget of one task (complex elements: 176 KB c128;
                                    # the task blob -- records + staged tables -- rides beside it)
SMALL_TASK_STEPS = 32               # single eval: steps per task (parallel subtrees run on
                                    # different CTAs; a batch runs one task chain per eval)
SMALL_MERGE_STEPS = 6               # a sibling subtree this short is absorbed, not exported
SMALL_MIN_STEPS = 8                 # fewer small steps than this: not worth the launch


SMALL_STAGE_WORDS = 6144             # a task's offset tables up to this size ride in shared memory
SP_HDR, SP_COPY_WORDS, SP_STEP_WORDS = 8, 8, 8


@dataclass
class SmallProgram:
    """Every early tiny step of a plan as ONE persistent launch.  Per task one BLOB of
    32-bit words (what the CTA copies into shared memory in a single coalesced pass):
    header [n_dep, n_imp, n_exp, n_steps, staged, 0, 0, 0] | dependency task ids |
    copy records (imports, then exports; 8 words: space, scratch offset, n, 0,
    global offset lo/hi, eval stride lo/hi) | step records (8 words: n_out, n_red,
    a_off, b_off, c_off, rsplit, out_tab, red_tab) | the task's offset tables when they
    are small enough to be staged (then out_tab / red_tab are word offsets into this
    part, else row offsets into the global tables)."""
    blob: np.ndarray                # uint32
    blob_off: np.ndarray            # int32 [n_tasks + 1] word offsets
    out_tab: np.ndarray             # uint32 [n, 2]: (oa | ob << 16, oc)  (unstaged tasks)
    red_tab: np.ndarray             # uint32: ra | rb << 16
    task_scratch: np.ndarray        # int32 [n_tasks]: scratch elements of each task
    step_ids: List[int]             # indices into plan.steps, in execution order
    task_of: Dict[int, int]         # plan step -> task

    @property
    def n_tasks(self) -> int:
        return len(self.blob_off) - 1

    def smem_bytes(self, esize: int) -> int:
        """Dynamic shared memory of the launch: per task its blob (16-byte aligned),
        then its scratch; the largest task decides."""
        words = np.diff(self.blob_off).astype(np.int64)
        return int((((words * 4 + 15) // 16) * 16 + self.task_scratch.astype(np.int64) * esize).max())


def _mode_offsets(modes: List[List[int]]):
    """(oa, ob, oc) of every index tuple of a mode list, C order (last fastest)."""
    oa = np.zeros(1, dtype=np.int64)
    ob, oc = oa.copy(), oa.copy()
    for d, sa, sb, sc in modes:
        idx = np.arange(d, dtype=np.int64)
        oa = (oa[:, None] + idx[None, :] * sa).reshape(-1)
        ob = (ob[:, None] + idx[None, :] * sb).reshape(-1)
        oc = (oc[:, None] + idx[None, :] * sc).reshape(-1)
    return oa, ob, oc


class _ScratchOverflow(Exception):
    pass


def _build_small_program(steps: List[Step], batch: int, arena_elems: int,
                         ws_top: int) -> Tuple[Optional[SmallProgram], int]:
    """``_build_small_program_once`` with the scratch budget lowered until the real
    shared-memory layout of every task (first-fit allocation fragments) fits an SM."""
    budget = SMALL_SCRATCH_ELEMS
    saved = [(st.c.offset, st.grouped) for st in steps]
    while True:
        try:
            return _build_small_program_once(steps, batch, arena_elems, ws_top, budget)
        except _ScratchOverflow:
            for st, (off, grouped) in zip(steps, saved):
                st.c.offset, st.grouped = off, grouped
            budget = int(budget * 0.8)
            if budget < 1024:
                return None, ws_top


def _build_small_program_once(steps: List[Step], batch: int, arena_elems: int, ws_top: int,
                              budget: int) -> Tuple[Optional[SmallProgram], int]:
    """Pick the EARLY tiny steps (operands are leaves or other early tiny steps), cut
    them into tasks, lay every task out in shared memory and tabulate the operand
    offsets of every output / reduction index (plan-time constants, shared by all
    evals of a batch and by every replay)."""
    producer = {id(st.c): k for k, st in enumerate(steps)}
    consumer: Dict[int, int] = {}
    for k, st in enumerate(steps):
        for lay in (st.a, st.b):
            p = producer.get(id(lay))
            if p is not None:
                consumer[p] = k
    early = [False] * len(steps)
    for k, st in enumerate(steps):
        if st.final or st.per_slice or st.a_slice or st.b_slice or st.small_modes is None:
            continue
        if max(st.a.size, st.b.size, st.c.size) > SMALL_ELEMS or \
                st.a.size + st.b.size + st.c.size > budget:     # all three sit in the scratch
            continue
        outm, redm = st.small_modes
        n_out = int(np.prod([m[0] for m in outm], dtype=np.int64)) if outm else 1
        n_red = int(np.prod([m[0] for m in redm], dtype=np.int64)) if redm else 1
        if n_out * n_red > SMALL_WORK or n_out > SMALL_ELEMS:
            continue
        ok = True
        for lay in (st.a, st.b):
            p = producer.get(id(lay))
            if (p is None and lay.space != ARENA) or (p is not None and not early[p]):
                ok = False
        early[k] = ok
    ids = [k for k in range(len(steps)) if early[k]]
    if len(ids) < SMALL_MIN_STEPS:
        return None, ws_top

    # ---- tasks: walk the steps in (topological) plan order; a step continues the task
    # of its larger operand subtree; the other operand's task is absorbed when short
    # (or whenever it fits, for a batch: parallelism comes from the evals), otherwise it
    # stays a task of its own (runs in parallel, result handed over in global memory).
    # Scratch need of a task = what it imports (resident throughout) + the peak of its
    # live intermediates (an operand dies at its single consumer).
    max_steps = SMALL_TASK_STEPS if batch == 1 else 1 << 30
    merge_steps = SMALL_MERGE_STEPS if batch == 1 else 1 << 30
    task_of: Dict[int, int] = {}
    members: Dict[int, List[int]] = {}
    resident: Dict[int, int] = {}               # imports of the task (elements)
    live: Dict[int, int] = {}                   # live intermediates right now
    peak: Dict[int, int] = {}                   # peak of resident + live

    def leaf_need(st):
        seen, tot = set(), 0
        for lay in (st.a, st.b):
            if id(lay) not in producer and id(lay) not in seen:
                seen.add(id(lay))
                tot += lay.size
        return tot
    nxt_task = 0
    for k in ids:
        st = steps[k]
        ops = {}
        for lay in (st.a, st.b):
            p = producer.get(id(lay))
            if p is not None:
                ops[task_of[p]] = lay.size
        order = sorted(ops, key=lambda t: -len(members[t]))
        home = None
        if order:
            t0 = order[0]
            if len(members[t0]) < max_steps and \
                    resident[t0] + live[t0] + st.c.size + leaf_need(st) <= budget:
                home = t0
        if home is None:
            home = nxt_task
            nxt_task += 1
            members[home], resident[home], live[home], peak[home] = [], 0, 0, 0
        resident[home] += leaf_need(st)
        for t in order:
            if t == home:
                continue
            fits = max(peak[t] + resident[home] + live[home],
                       resident[t] + live[t] + resident[home] + live[home] + st.c.size) \
                <= budget
            if len(members[t]) <= merge_steps and len(members[home]) + len(members[t]) < max_steps \
                    and fits:
                for q in members[t]:
                    task_of[q] = home
                peak[home] = max(peak[home], peak[t] + resident[home] + live[home])
                members[home] = sorted(members[home] + members.pop(t))
                resident[home] += resident.pop(t)
                live[home] += live.pop(t)
                peak.pop(t)
            else:
                resident[home] += ops[t]          # imported hand-over
        members[home].append(k)
        task_of[k] = home
        live[home] += st.c.size
        peak[home] = max(peak[home], resident[home] + live[home])
        for lay in (st.a, st.b):                  # in-task operands die here
            p = producer.get(id(lay))
            if p is not None and task_of[p] == home:
                live[home] -= lay.size
    # renumber tasks in order of their LAST step: a task's dependencies then have
    # smaller numbers (tickets are handed out in task order: no waiting on a later one)
    order = sorted(members, key=lambda t: members[t][-1])
    renum = {t: i for i, t in enumerate(order)}
    task_of = {k: renum[t] for k, t in task_of.items()}
    members = {renum[t]: ks for t, ks in members.items()}
    n_tasks = len(members)
    # what a task hands over (to another task or to a later launch) gets a private,
    # never recycled workspace buffer: the program runs before every other launch
    for k in ids:
        st = steps[k]
        c = consumer.get(k)
        if (c is None or task_of.get(c) != task_of[k]) and st.c.space == WS:
            st.c.offset = int(ws_top)
            ws_top += st.c.size * batch

    out_rows: List[np.ndarray] = []
    red_rows: List[np.ndarray] = []
    out_cache: Dict[bytes, int] = {}
    red_cache: Dict[bytes, int] = {}
    n_out_tab = n_red_tab = 0
    task_scratch: List[int] = []
    step_ids: List[int] = []
    blobs: List[np.ndarray] = []
    blob_off = [0]

    def eval_stride(lay):
        if batch == 1:
            return 0
        return arena_elems if lay.space == ARENA else lay.size

    def copy_words(space, smem_off, glob_off, stride, n):
        return [space, smem_off, n, 0, glob_off & 0xffffffff, glob_off >> 32,
                stride & 0xffffffff, stride >> 32]
    for t in range(n_tasks):
        ks = members[t]
        inside = set(ks)
        alloc = _Allocator()
        where: Dict[int, int] = {}               # id(layout) -> scratch element offset
        imports, exports, deps = [], [], []
        for k in ks:                             # imports first: resident for the whole task
            for lay in (steps[k].a, steps[k].b):
                if id(lay) in where:
                    continue
                p = producer.get(id(lay))
                if p is None or p not in inside:
                    where[id(lay)] = alloc.alloc(lay.size)
                    imports.append(copy_words(lay.space, where[id(lay)], lay.offset,
                                              eval_stride(lay), lay.size))
                    if p is not None:
                        deps.append(task_of[p])
        tabs = []                                # (otab, rtab) per step
        rows = []
        for k in ks:
            st = steps[k]
            outm, redm = st.small_modes
            c_off = alloc.alloc(st.c.size)
            where[id(st.c)] = c_off
            oa, ob, oc = _mode_offsets(outm)
            ra, rb, _ = _mode_offsets(redm)
            assert oa.max() + ra.max() < st.a.size and ob.max() + rb.max() < st.b.size
            otab = np.stack([(oa | (ob << 16)), oc], axis=1).astype(np.uint32)
            rtab = (ra | (rb << 16)).astype(np.uint32)
            n_out, n_red = len(otab), len(rtab)
            rsplit = 1                            # lanes sharing one output's reduction
            while rsplit < 32 and n_out * rsplit * 2 <= 256 and n_red >= 4 * rsplit:
                rsplit *= 2
            rows.append([n_out, n_red, where[id(st.a)], where[id(st.b)], c_off, rsplit, 0, 0])
            tabs.append((otab, rtab))
            step_ids.append(k)
            if consumer.get(k) not in inside:     # leaves the task: a root or a hand-over
                exports.append(copy_words(st.c.space, c_off, st.c.offset, eval_stride(st.c),
                                          st.c.size))
            for lay in (st.a, st.b):              # single consumer: operands die here
                p = producer.get(id(lay))
                if p is not None and p in inside and id(lay) in where:
                    alloc.release(where.pop(id(lay)), lay.size)
        deps = sorted(set(deps))
        assert all(d < t for d in deps), "task order is not topological"
        tab_words = sum(2 * len(o) + len(r) + (len(r) & 1) for o, r in tabs)
        staged = tab_words <= SMALL_STAGE_WORDS
        head_words = SP_HDR + len(deps) + (len(deps) & 1) + \
            SP_COPY_WORDS * (len(imports) + len(exports)) + SP_STEP_WORDS * len(rows)
        tail: List[np.ndarray] = []
        at = head_words                           # word offset of the staged tables in the blob
        for row, (otab, rtab) in zip(rows, tabs):
            if staged:
                row[6], row[7] = at, at + 2 * len(otab)
                tail.append(otab.reshape(-1))
                tail.append(rtab)
                if len(rtab) & 1:
                    tail.append(np.zeros(1, dtype=np.uint32))
                at += 2 * len(otab) + len(rtab) + (len(rtab) & 1)
            else:
                key = otab.tobytes()
                o_at = out_cache.get(key)
                if o_at is None:
                    o_at = out_cache[key] = n_out_tab
                    out_rows.append(otab)
                    n_out_tab += len(otab)
                key = rtab.tobytes()
                r_at = red_cache.get(key)
                if r_at is None:
                    r_at = red_cache[key] = n_red_tab
                    red_rows.append(rtab)
                    n_red_tab += len(rtab)
                row[6], row[7] = o_at, r_at
        head = [len(deps), len(imports), len(exports), len(rows), int(staged), 0, 0, 0]
        head += deps + ([0] if len(deps) & 1 else [])
        for c in imports + exports:
            head += c
        for row in rows:
            head += row
        assert len(head) == head_words
        blob = np.concatenate([np.array(head, dtype=np.int64).astype(np.uint32)] + tail)
        blobs.append(blob)
        blob_off.append(blob_off[-1] + len(blob))
        task_scratch.append(int(alloc.top))
        if ((len(blob) * 4 + 15) // 16) * 16 + alloc.top * 16 > 226 * 1024:
            raise _ScratchOverflow()
    return SmallProgram(
        np.concatenate(blobs), np.array(blob_off, dtype=np.int32),
        np.concatenate(out_rows).reshape(-1, 2) if out_rows else np.zeros((1, 2), dtype=np.uint32),
        np.concatenate(red_rows).reshape(-1) if red_rows else np.zeros(1, dtype=np.uint32),
        np.array(task_scratch, dtype=np.int32), step_ids, task_of), ws_top


def plan_bound_seconds(plan: Plan, hbm_gbs: float, pipe_tflops: float) -> float:
    """Roofline lower bound of one evaluation of ``plan``: per
    step max(algorithmic bytes / HBM bandwidth, 8 * MACs / math-pipe peak),
    summed over the steps (per-slice steps x nslices)."""
    total = 0.0
    for st in plan.steps:
        t = max(st.bytes / (hbm_gbs * 1e9), 8.0 * st.macs / (pipe_tflops * 1e12))
        total += t * (plan.nslices if st.per_slice else 1)
    return total


def _reallocate_for_slices(steps: List[Step], batch: int) -> None:
    """Invariant intermediates get a private region (never recycled by
    per-slice tensors); per-slice tensors share a second, recycled region."""
    inv = _Allocator()
    for st in steps:
        if st.c.space == WS and not st.per_slice:
            st.c.offset = inv.alloc(st.c.size * batch)
    base = inv.top
    per = _Allocator()
    per_slice_ids = {id(st.c) for st in steps if st.per_slice}
    for st in steps:
        if not st.per_slice:
            continue
        if st.c.space == WS:
            st.c.offset = base + per.alloc(st.c.size * batch)
        for lay in (st.a, st.b):
            if lay.space == WS and id(lay) in per_slice_ids:
                per.release(lay.offset - base, lay.size * batch)


@dataclass
class GivenTree:
    """An explicit contraction tree for ``B200Backend(hyperoptimiser=...)``: SSA pair
    order over the network's leaves plus the indices to slice (network index labels)."""
    ssa_path: List[Tuple[int, ...]]
    sliced_inds: Tuple[int, ...] = ()


def info_from_ssa_path(net: Network, ssa_path: Sequence[Sequence[int]],
                       sliced_inds: Sequence[int] = ()) -> PathInfo:
    """A contraction tree found elsewhere -- cotengra's ``tree.get_ssa_path()`` and
    ``tree.sliced_inds`` for the same tensors in the same order (the reference hands its
    network to a cotengra ``HyperOptimizer``, optyx/core/backends.py:436-455, 539-545) --
    as a ``PathInfo``.  SSA form: leaves are ids ``0..n-1``, the k-th entry creates id
    ``n + k``; an entry with more than two ids is contracted left to right.  Leftover
    components (a path that stops early) are multiplied together at the end."""
    n = len(net.leaves)
    inputs = [frozenset(l.inds) for l in net.leaves]
    output = frozenset(net.output_inds)
    sliced = tuple(dict.fromkeys(int(q) for q in sliced_inds))
    for q in sliced:
        if q not in net.dims:
            raise ValueError(f"sliced index {q} is not an index of the network")
        if q in output:
            raise ValueError(f"sliced index {q} is an open index: only summed indices can be sliced")
    live = set(range(n))
    path: List[Tuple[int, int]] = []
    rename: Dict[int, int] = {}          # caller's SSA id -> ours (differs after a >2-way entry)
    nxt_theirs, nxt_ours = n, n
    for entry in ssa_path:
        ids = [rename.get(int(i), int(i)) for i in entry]
        if len(ids) < 2 or len(set(ids)) != len(ids) or any(i not in live for i in ids):
            raise ValueError(f"invalid SSA path entry {tuple(entry)}: ids must be distinct, "
                             "live tensors")
        cur = ids[0]
        for other in ids[1:]:
            path.append((cur, other))
            live.discard(cur)
            live.discard(other)
            cur = nxt_ours
            live.add(cur)
            nxt_ours += 1
        rename[nxt_theirs] = cur
        nxt_theirs += 1
    rest = sorted(live)
    while len(rest) > 1:
        path.append((rest[0], rest[1]))
        rest = rest[2:] + [nxt_ours]
        nxt_ours += 1
    return analyse_path(inputs, net.dims, output, path, sliced)


def plan_from_ssa_path(net: Network, ssa_path: Sequence[Sequence[int]],
                       sliced_inds: Sequence[int] = (), dtype=np.complex128, batch: int = 1,
                       gemm_threshold: Tuple = (32, 16, 8, 4.0)) -> Plan:
    """Executable plan of a GIVEN contraction tree (see ``info_from_ssa_path``):
    layouts, kernel choice per step, slicing by pointer offsets, workspace."""
    info = info_from_ssa_path(net, ssa_path, sliced_inds) if len(net.leaves) >= 2 \
        else PathInfo([], (), 0.0, 0.0, 1, [])
    return compile_plan(net, info, dtype=dtype, batch=batch, gemm_threshold=gemm_threshold)


def plan_network(net: Network, dtype=np.complex128, batch: int = 1, trials: int = 8,
                 seed: int = 0, target_size: Optional[float] = None,
                 gemm_threshold: Tuple = (32, 16, 8, 4.0), min_slices: int = 1) -> Plan:
    """Path search (+ slicing to ``target_size`` elements / ``min_slices`` slices) +
    compilation."""
    inputs = [frozenset(l.inds) for l in net.leaves]
    output = frozenset(net.output_inds)
    if len(inputs) >= 2:
        info = find_path(inputs, net.dims, output, trials=trials, seed=seed,
                         target_size=target_size, min_slices=min_slices)
    else:
        info = PathInfo([], (), 0.0, 0.0, 1, [])
    return compile_plan(net, info, dtype=dtype, batch=batch, gemm_threshold=gemm_threshold)


# --------------------------------------------------------------------------
# native (C++) path search: csrc/planner_native.cu behind the C ABI
# --------------------------------------------------------------------------
class NativeNetwork:
    """A network in the CSR form ``b2_plan_greedy`` / ``b2_plan_cost`` take: compact
    index ids, extents, open-index flags.  Built once per search."""

    def __init__(self, inputs: Sequence[frozenset], dims: Dict[int, int], output: frozenset):
        from . import runtime as rt
        self.lib = rt.load()
        ids = sorted({q for s in inputs for q in s} | set(output))
        self.ids = ids
        self.col = {q: k for k, q in enumerate(ids)}
        self.n = len(inputs)
        ptr = np.zeros(self.n + 1, dtype=np.int32)
        flat: List[int] = []
        for k, s in enumerate(inputs):
            flat.extend(self.col[q] for q in s)
            ptr[k + 1] = len(flat)
        self.ptr = ptr
        self.flat = np.array(flat, dtype=np.int32) if flat else np.zeros(1, dtype=np.int32)
        self.dims = np.array([dims[q] for q in ids], dtype=np.int32) if ids else np.zeros(1, dtype=np.int32)
        out = np.zeros(max(1, len(ids)), dtype=np.uint8)
        for q in output:
            out[self.col[q]] = 1
        self.is_output = out

    def greedy(self, seed: int = 0, temperature: float = 0.0, costmod: float = 1.0) -> np.ndarray:
        """SSA pair order as an int32 array [n - 1, 2]."""
        from . import runtime as rt
        path = np.zeros(max(1, 2 * (self.n - 1)), dtype=np.int32)
        rt.check(self.lib.b2_plan_greedy(
            self.n, self.ptr.ctypes.data, self.flat.ctypes.data, len(self.ids), self.dims.ctypes.data,
            self.is_output.ctypes.data, seed, temperature, costmod, path.ctypes.data), "b2_plan_greedy")
        return path[:2 * (self.n - 1)].reshape(-1, 2)

    def reconfigure(self, path: np.ndarray, sliced: Sequence[int] = (), max_leaves: int = 8,
                    sweeps: int = 4) -> np.ndarray:
        """Subtree reconfiguration (``b2_plan_reconfigure``) until nothing improves."""
        from . import runtime as rt
        import ctypes as C
        path = np.ascontiguousarray(path, dtype=np.int32).copy()
        flags = np.zeros(max(1, len(self.ids)), dtype=np.uint8)
        for q in sliced:
            flags[self.col[q]] = 1
        improved = C.c_int32(0)
        for _ in range(sweeps):
            rt.check(self.lib.b2_plan_reconfigure(
                self.n, self.ptr.ctypes.data, self.flat.ctypes.data, len(self.ids),
                self.dims.ctypes.data, self.is_output.ctypes.data, flags.ctypes.data, max_leaves,
                path.ctypes.data, C.byref(improved)), "b2_plan_reconfigure")
            if improved.value == 0:
                break
        return path

    def cost(self, path: np.ndarray, sliced: Sequence[int] = (), want_sizes: bool = False):
        """(macs per slice, peak elements, nslices, time estimate[, log2 step sizes])."""
        from . import runtime as rt
        path = np.ascontiguousarray(path, dtype=np.int32)
        flags = np.zeros(max(1, len(self.ids)), dtype=np.uint8)
        for q in sliced:
            flags[self.col[q]] = 1
        out = np.zeros(4, dtype=np.float64)
        sizes = np.zeros(max(1, self.n - 1), dtype=np.float64) if want_sizes else None
        rt.check(self.lib.b2_plan_cost(
            self.n, self.ptr.ctypes.data, self.flat.ctypes.data, len(self.ids), self.dims.ctypes.data,
            self.is_output.ctypes.data, path.ctypes.data, flags.ctypes.data, out.ctypes.data,
            sizes.ctypes.data if want_sizes else None), "b2_plan_cost")
        res = (float(out[0]), float(out[1]), int(round(out[2])), float(out[3]))
        return res + (sizes,) if want_sizes else res
