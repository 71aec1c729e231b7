"""Test infrastructure: random networks and a numpy *emulator of the kernel
descriptors* (so planner/layout/slicing logic is checked on CPU; the CUDA
kernels themselves are checked by the ``-m gpu`` tests against the oracle)."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from optyx_b200.core.network import (K_ADD, K_DENSE, K_EMBED, K_ONEHOT, K_SUM, K_VEC,  # noqa: E402
                                     Leaf, Network)
from optyx_b200.planner import ARENA, OUT, WS, Plan  # noqa: E402
from optyx_b200 import runtime as rt  # noqa: E402
from oracle import leaves as oracle_leaves  # noqa: E402


def random_network(rng: np.random.Generator, n_leaves=8, n_inds=10, max_rank=4, dims=(2, 3, 4),
                   n_out=2, hyper=True, repeat_in_leaf=False) -> Network:
    """Random hyper-network of DENSE / W / ADD / VEC / ONEHOT / EMBED leaves."""
    dmap = {i: int(rng.choice(dims)) for i in range(n_inds)}
    leaves = []
    used = set()
    for k in range(n_leaves):
        rank = int(rng.integers(1, min(max_rank, n_inds) + 1))
        if repeat_in_leaf and k == 0 and rank >= 2:
            base = list(rng.choice(n_inds, size=rank - 1, replace=False))
            inds = tuple(int(i) for i in base + [base[0]])
        else:
            inds = tuple(int(i) for i in rng.choice(n_inds, size=rank, replace=False))
        used.update(inds)
        shape = tuple(dmap[i] for i in inds)
        kind = int(rng.choice([K_DENSE, K_DENSE, K_SUM, K_ADD, K_VEC, K_ONEHOT, K_EMBED]))
        params, ipar = None, 0
        if kind == K_VEC and rank != 1:
            kind = K_DENSE
        if kind == K_ONEHOT and rank != 1:
            kind = K_DENSE
        if kind == K_EMBED and rank != 2:
            kind = K_DENSE
        if kind in (K_SUM, K_ADD) and rank < 2:
            kind = K_DENSE
        if kind in (K_DENSE, K_VEC):
            params = rng.standard_normal(int(np.prod(shape))) + \
                1j * rng.standard_normal(int(np.prod(shape)))
        if kind == K_ONEHOT:
            ipar = int(rng.integers(0, shape[0] + 1))
        leaves.append(Leaf(kind, shape, inds, ipar, params, f"leaf{k}"))
    if not hyper:
        pass
    used = sorted(used)
    out = tuple(int(i) for i in rng.choice(used, size=min(n_out, len(used)), replace=False))
    dims_used = {i: dmap[i] for i in used}
    net = Network(leaves, dims_used, out, scalar=complex(0.7, -0.2), n_dom=0)
    net.validate()
    return net


# ---- descriptor emulation ---------------------------------------------------
def _offsets(dimlist, strides_list):
    """All offsets of a mode list (C order over the list, last fastest)."""
    offs = [np.zeros(1, dtype=np.int64) for _ in strides_list]
    for d, *ss in zip(dimlist, *strides_list):
        idx = np.arange(d, dtype=np.int64)
        offs = [(o[:, None] + idx[None, :] * s).reshape(-1) for o, s in zip(offs, ss)]
    return offs


def emulate_direct(desc: rt.Modes, A, a_off, B, b_off, Cbuf, c_off, alpha, accumulate):
    no, nr = desc.n_out, desc.n_red
    dim = list(desc.dim[:no + nr])
    sa, sb, sc = list(desc.sa[:no + nr]), list(desc.sb[:no + nr]), list(desc.sc[:no + nr])
    oa, ob, oc = _offsets(dim[:no], [sa[:no], sb[:no], sc[:no]])
    ra, rb = _offsets(dim[no:], [sa[no:], sb[no:]])
    acc = (A[a_off + oa[:, None] + ra[None, :]] * B[b_off + ob[:, None] + rb[None, :]]).sum(axis=1)
    res = alpha * acc
    if accumulate:
        Cbuf[c_off + oc] += res
    else:
        Cbuf[c_off + oc] = res


def emulate_gemm(desc: rt.GemmModes, A, a_off, B, b_off, Cbuf, c_off, alpha, accumulate):
    nb, nm, nn, nk = desc.n_batch, desc.n_m, desc.n_n, desc.n_k
    tot = nb + nm + nn + nk
    dim = list(desc.dim[:tot])
    sa, sb, sc = list(desc.sa[:tot]), list(desc.sb[:tot]), list(desc.sc[:tot])

    def grp(lo, hi):
        # group mode 0 is fastest -> reverse to reuse the C-order helper
        sl = slice(lo, hi)
        return _offsets(dim[sl][::-1], [sa[sl][::-1], sb[sl][::-1], sc[sl][::-1]])
    ba, bb, bc = grp(0, nb)
    ma, _, mc = grp(nb, nb + nm)
    _, nb_, nc = grp(nb + nm, nb + nm + nn)
    ka, kb, _ = grp(nb + nm + nn, tot)
    for e in range(len(ba)):
        Am = A[a_off + ba[e] + ma[:, None] + ka[None, :]]
        Bm = B[b_off + bb[e] + kb[:, None] + nb_[None, :]]
        res = alpha * (Am @ Bm)
        idx = c_off + bc[e] + mc[:, None] + nc[None, :]
        if accumulate:
            Cbuf[idx] += res
        else:
            Cbuf[idx] = res


def emulate_small_program(sp, bufs, batch: int) -> None:
    """numpy model of csrc/contract_small.cu, parsing the same task BLOBS the kernel
    copies into shared memory: (eval, task) items in ticket order, every task on a
    private NaN-filled scratch -- imports, the tabulated steps, exports; a dependency
    that has not run yet, or a read of something never written, fails."""
    done = set()
    ot_all, rt_all = sp.out_tab.astype(np.int64), sp.red_tab.astype(np.int64)
    assert sp.smem_bytes(16) <= 226 * 1024
    for ev in range(batch):
        for t in range(sp.n_tasks):
            H = sp.blob[sp.blob_off[t]:sp.blob_off[t + 1]].astype(np.int64)
            n_dep, n_imp, n_exp, n_steps, staged = (int(x) for x in H[:5])
            at = 8
            deps = H[at:at + n_dep]
            at += n_dep + (n_dep & 1)
            copies = H[at:at + 8 * (n_imp + n_exp)].reshape(-1, 8)
            at += 8 * (n_imp + n_exp)
            steps = H[at:at + 8 * n_steps].reshape(-1, 8)
            for d in deps:
                assert (ev, int(d)) in done and int(d) < t, "dependency runs later"
            S = np.full(int(sp.task_scratch[t]), np.nan + 0j, dtype=complex)

            def glob(c):
                return int(c[4] | (c[5] << 32)) + ev * int(c[6] | (c[7] << 32))
            for c in copies[:n_imp]:
                chunk = bufs[int(c[0])][glob(c):glob(c) + int(c[2])]
                assert not np.isnan(chunk).any(), "import of an unwritten buffer"
                S[int(c[1]):int(c[1]) + int(c[2])] = chunk
            for st in steps:
                n_out, n_red, a_off, b_off, c_off, rsplit, o_at, r_at = (int(x) for x in st)
                if staged:
                    assert o_at % 2 == 0
                    ot = H[o_at:o_at + 2 * n_out].reshape(-1, 2)
                    rt = H[r_at:r_at + n_red]
                else:
                    ot, rt = ot_all[o_at:o_at + n_out], rt_all[r_at:r_at + n_red]
                oa, ob, oc = ot[:, 0] & 0xffff, ot[:, 0] >> 16, ot[:, 1]
                ra, rb = rt & 0xffff, rt >> 16
                assert rsplit in (1, 2, 4, 8, 16, 32)
                vals = (S[a_off + oa[:, None] + ra[None, :]]
                        * S[b_off + ob[:, None] + rb[None, :]]).sum(axis=1)
                assert not np.isnan(vals).any(), "step reads scratch nobody wrote"
                S[c_off + oc] = vals
            for c in copies[n_imp:]:
                bufs[int(c[0])][glob(c):glob(c) + int(c[2])] = S[int(c[1]):int(c[1]) + int(c[2])]
            done.add((ev, t))


def emulate_plan(net: Network, plan: Plan, nets=None, slice_range=None) -> np.ndarray:
    """Run a plan on numpy buffers with the emulated kernels."""
    nets = nets or [net]
    b = plan.batch
    arena = np.zeros(b * plan.arena_elems, dtype=complex)
    for e, ne in enumerate(nets):
        for off, leaf in zip(plan.leaf_offsets, ne.leaves):
            arr = oracle_leaves.build_leaf(leaf.kind, leaf.shape, leaf.ipar, leaf.params)
            arena[e * plan.arena_elems + off: e * plan.arena_elems + off + leaf.size] = \
                arr.reshape(-1)
        arena[e * plan.arena_elems + plan.unit_offset] = 1.0
    ws = np.full(max(1, plan.ws_elems), np.nan + 0j, dtype=complex)
    out = np.full(b * plan.out_elems, np.nan + 0j, dtype=complex)
    if plan.needs_zero_out:
        out[:] = 0
    bufs = {ARENA: arena, WS: ws, OUT: out}

    def launch(st, vals, alpha, accumulate):
        a_off = st.a.offset + sum(vals[p] * s for p, s in st.a_slice)
        b_off = st.b.offset + sum(vals[p] * s for p, s in st.b_slice)
        fn = emulate_gemm if st.kernel == "gemm" else emulate_direct
        fn(st.desc, bufs[st.a.space], a_off, bufs[st.b.space], b_off, bufs[st.c.space],
           st.c.offset, alpha, accumulate)

    scalar = nets[0].scalar
    zeros = [0] * len(plan.sliced)
    if plan.small is not None:            # the persistent small-step launch runs first
        emulate_small_program(plan.small, bufs, b)
    for st in plan.steps:
        if not st.per_slice and not st.grouped:
            launch(st, zeros, scalar if st.final else 1.0, 0)
    if plan.sliced:
        s0, s1 = slice_range or (0, plan.nslices)
        for sid in range(s0, s1):
            vals, rem = [0] * len(plan.sliced), sid
            for k in range(len(plan.sliced) - 1, -1, -1):
                vals[k] = rem % plan.sliced_dims[k]
                rem //= plan.sliced_dims[k]
            for st in plan.steps:
                if st.per_slice:
                    launch(st, vals, scalar if st.final else 1.0, 1 if st.final else 0)
    return out.reshape((b,) + tuple(plan.out_shape))
