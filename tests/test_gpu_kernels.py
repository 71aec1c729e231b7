"""-m gpu: every CUDA kernel, called through the C ABI, against the oracle."""
import ctypes as C

import numpy as np
import pytest

from helpers import emulate_plan, random_network
from oracle import evalresult as oracle_eval
from oracle import leaves as oracle_leaves
from oracle.contract import contract_network
from optyx_b200 import runtime as rt
from optyx_b200.core.network import (K_ADD, K_DENSE, K_DIV, K_DUALRAIL, K_EMBED, K_MOD2, K_MUL,
                                     K_ONEHOT, K_ONES, K_PTD, K_SUM, K_VEC, Leaf, Network)

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    return torch


def _net_of_leaves(specs, rng):
    leaves, dims, nxt = [], {}, 0
    for kind, shape, ipar in specs:
        inds = tuple(range(nxt, nxt + len(shape)))
        nxt += len(shape)
        for i, d in zip(inds, shape):
            dims[i] = d
        params = None
        if kind in (K_VEC, K_DENSE):
            n = int(np.prod(shape))
            params = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        leaves.append(Leaf(kind, tuple(shape), inds, ipar, params))
    return Network(leaves, dims, (), 1.0)


LEAF_SPECS = [
    (K_ONEHOT, (4,), 2), (K_ONEHOT, (3,), 5), (K_ONES, (7,), 0), (K_VEC, (7,), 0),
    (K_DENSE, (2, 3, 4), 0), (K_SUM, (4, 4, 4), 0), (K_SUM, (7, 7, 7), 0), (K_SUM, (7, 3, 5), 0),
    (K_SUM, (10, 4, 4, 4), 0), (K_SUM, (13, 7, 7), 0), (K_SUM, (5, 5), 0), (K_ADD, (7, 4, 4), 0),
    (K_ADD, (5, 3, 3, 3), 0), (K_MUL, (10, 4, 4), 0), (K_DIV, (10, 10, 4), 0),
    (K_MOD2, (2, 7), 0), (K_DUALRAIL, (2, 2, 2), 0), (K_DUALRAIL, (2, 3, 3), 0),
    (K_PTD, (2, 6), 0), (K_EMBED, (3, 7), 0), (K_EMBED, (7, 2), 0), (K_SUM, (20, 10, 10, 10), 0),
]


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("batch", [1, 3])
def test_leaf_build_matches_oracle(dtype, batch):
    from optyx_b200.executor import Session
    from optyx_b200.planner import plan_network
    rng = np.random.default_rng(1)
    nets = [_net_of_leaves(LEAF_SPECS, np.random.default_rng(10 + e)) for e in range(batch)]
    plan = plan_network(nets[0], dtype=dtype, batch=batch)
    sess = Session(plan)
    sess.set_leaves(nets)
    sess.build_leaves()
    arena = sess.arena.cpu().numpy().reshape(batch, plan.arena_elems)
    for e, net in enumerate(nets):
        for off, leaf in zip(plan.leaf_offsets, net.leaves):
            want = oracle_leaves.build_leaf(leaf.kind, leaf.shape, leaf.ipar, leaf.params)
            got = arena[e, off:off + leaf.size].reshape(leaf.shape)
            # exact structure (non-zero pattern) ...
            assert np.array_equal(got != 0, want != 0), (leaf.kind, leaf.shape)
            # ... and values: host params pass through unchanged, sqrt(multinomial)
            # within 1 ulp of the reference's `** 0.5`
            tol = 4e-16 if dtype == np.complex128 else 2e-7
            assert np.allclose(got, want, rtol=tol, atol=0), (leaf.kind, leaf.shape)
        assert arena[e, plan.unit_offset] == 1.0


@pytest.mark.parametrize("seed", range(40))
def test_random_networks_match_oracle(seed):
    from optyx_b200.executor import Session
    from optyx_b200.planner import plan_network
    rng = np.random.default_rng(seed)
    net = random_network(rng, n_leaves=int(rng.integers(1, 10)), n_inds=int(rng.integers(3, 12)),
                         n_out=int(rng.integers(0, 4)), repeat_in_leaf=seed % 3 == 0)
    if seed % 5 == 0:
        net.output_inds = net.output_inds + net.output_inds[:1]
    want = contract_network(net)
    for thr in [(16, 16, 4), (1, 1, 1)]:
        for tgt in [None, 8]:
            plan = plan_network(net, target_size=tgt, gemm_threshold=thr, trials=3)
            got = Session(plan).run([net])[0].cpu().numpy()
            assert got.shape == want.shape
            scale = max(1.0, np.abs(want).max())
            assert np.abs(got - want).max() <= 1e-10 * scale, (seed, thr, tgt)


@pytest.mark.parametrize("seed", range(6))
def test_random_networks_complex64(seed):
    from optyx_b200.executor import Session
    from optyx_b200.planner import plan_network
    rng = np.random.default_rng(100 + seed)
    net = random_network(rng, n_leaves=7, n_inds=9, n_out=2)
    want = contract_network(net)
    plan = plan_network(net, dtype=np.complex64, target_size=16 if seed % 2 else None)
    got = Session(plan).run([net])[0].cpu().numpy()
    scale = max(1.0, np.abs(want).max())
    assert np.abs(got - want).max() <= 1e-5 * scale


@pytest.mark.parametrize("batch", [2, 5])
def test_batched_evals_share_one_plan(batch):
    from optyx_b200.executor import Session
    from optyx_b200.planner import plan_network
    nets = []
    for e in range(batch):
        rng = np.random.default_rng(7)          # same structure ...
        net = random_network(rng, n_leaves=8, n_inds=9, n_out=2)
        prng = np.random.default_rng(1000 + e)  # ... different values
        for leaf in net.leaves:
            if leaf.params is not None:
                leaf.params = prng.standard_normal(leaf.size) + 1j * prng.standard_normal(leaf.size)
        nets.append(net)
    for thr in [(16, 16, 4), (1, 1, 1)]:
        plan = plan_network(nets[0], batch=batch, gemm_threshold=thr)
        got = Session(plan).run(nets).cpu().numpy()
        for e, net in enumerate(nets):
            want = contract_network(net)
            assert np.abs(got[e] - want).max() <= 1e-10 * max(1.0, np.abs(want).max())


def _gemm_case(rng, mshape, nshape, kshape, bshape=(), dtype=np.complex128, accumulate=0):
    """A[b,m,k] x B[b,k,n] with every operand stored in a random axis order
    (``accumulate``: C starts from random values and the product is added to it)."""
    torch = _torch()
    lib = rt.load()
    names = {}
    groups = {"b": bshape, "m": mshape, "n": nshape, "k": kshape}
    for g, shp in groups.items():
        for i, d in enumerate(shp):
            names[f"{g}{i}"] = d
    def tensor(groups_in):
        axes = [f"{g}{i}" for g in groups_in for i in range(len(groups[g]))]
        perm = list(rng.permutation(len(axes)))
        axes = [axes[p] for p in perm]
        shape = [names[a] for a in axes]
        arr = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(dtype)
        strides, s = {}, 1
        for a, d in zip(reversed(axes), reversed(shape)):
            strides[a] = s
            s *= d
        return arr, axes, strides
    A, axA, sA = tensor("bmk")
    Bm, axB, sB = tensor("bkn")
    Cz, axC, sC = tensor("bmn")
    desc = rt.GemmModes()
    desc.n_batch, desc.n_m, desc.n_n, desc.n_k = len(bshape), len(mshape), len(nshape), len(kshape)
    p = 0
    for g in "bmnk":
        for i in range(len(groups[g])):
            a = f"{g}{i}"
            desc.dim[p] = names[a]
            desc.sa[p], desc.sb[p], desc.sc[p] = sA.get(a, 0), sB.get(a, 0), sC.get(a, 0)
            p += 1
    letters = {a: chr(ord("a") + i) for i, a in enumerate(names)}
    spec = "".join(letters[a] for a in axA) + "," + "".join(letters[a] for a in axB) + "->" + \
        "".join(letters[a] for a in axC)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(Bm).cuda()
    dC = torch.from_numpy(Cz.copy()).cuda() if accumulate else \
        torch.zeros(Cz.shape, dtype=dA.dtype, device="cuda")
    alpha = 0.5 - 0.25j
    stream = torch.cuda.current_stream().cuda_stream
    code = rt.B2_C128 if dtype == np.complex128 else rt.B2_C64
    rt.check(lib.b2_contract_gemm(C.byref(desc), dA.data_ptr(), dB.data_ptr(), dC.data_ptr(),
                                  alpha.real, alpha.imag, int(accumulate), code, stream))
    torch.cuda.synchronize()
    want = alpha * torch.einsum(spec, dA.to(torch.complex128), dB.to(torch.complex128))
    if accumulate:
        want = want + torch.from_numpy(Cz).cuda().to(torch.complex128)
    return dC.to(torch.complex128), want


def _direct_case(rng, out_shape, red_shape, dtype, accumulate):
    """C[out] (+)= alpha * sum_red A[..] B[..] through b2_contract_direct; every mode is
    given at random to A, B or both, operands stored in random axis orders."""
    torch = _torch()
    lib = rt.load()
    n_out, n_red = len(out_shape), len(red_shape)
    dims = list(out_shape) + list(red_shape)
    in_a, in_b = [], []
    for m in range(len(dims)):
        w = int(rng.integers(0, 3)) if m < n_out else int(rng.integers(0, 4))
        (in_a if w in (0, 2, 3) else in_b).append(m)
        if w in (2, 3) and m not in in_b:
            in_b.append(m)
    def operand(modes):
        order = [modes[i] for i in rng.permutation(len(modes))]
        shape = [dims[m] for m in order]
        arr = np.asarray(rng.standard_normal(shape) + 1j * rng.standard_normal(shape))
        strides, s = {}, 1
        for m, d in zip(reversed(order), reversed(shape)):
            strides[m] = s
            s *= d
        return arr.astype(dtype), order, strides
    A, oA, sA = operand(in_a)
    B, oB, sB = operand(in_b)
    md = rt.Modes()
    md.n_out, md.n_red = n_out, n_red
    s = 1
    for m in range(len(dims)):
        md.dim[m] = dims[m]
        md.sa[m], md.sb[m] = sA.get(m, 0), sB.get(m, 0)
    for m in reversed(range(n_out)):
        md.sc[m] = s
        s *= dims[m]
    letters = [chr(ord("a") + m) for m in range(len(dims))]
    spec = "".join(letters[m] for m in oA) + "," + "".join(letters[m] for m in oB) + "->" + \
        "".join(letters[:n_out])
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    c0 = np.asarray(rng.standard_normal(out_shape) + 1j * rng.standard_normal(out_shape)).astype(dtype)
    dC = torch.from_numpy(c0.copy()).cuda()
    alpha = 0.5 - 0.25j
    stream = torch.cuda.current_stream().cuda_stream
    dt = rt.B2_C128 if dtype == np.complex128 else rt.B2_C64
    rt.check(lib.b2_contract_direct(C.byref(md), dA.data_ptr(), dB.data_ptr(), dC.data_ptr(),
                                    alpha.real, alpha.imag, accumulate, dt, stream))
    kernel = lib.b2_last_kernel().decode()
    want = alpha * np.einsum(spec, A.astype(np.complex128), B.astype(np.complex128))
    if accumulate:
        want = want + c0
    return dC.cpu().numpy(), want, kernel


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("accumulate", [0, 1])
@pytest.mark.parametrize("case", [
    ((7, 7), (7, 7, 7, 7, 7)),                  # reduction-dominated (split-K kernel)
    ((), (2,) * 14),                            # scalar output, long reduction
    ((3,), (5, 1000)),
    ((2,) * 16, (2, 2)),                        # big x tiny
    ((7, 7, 7, 7), (7,)),
    ((4, 4, 4, 4, 4, 4), (4, 4)),
    ((2,) * 20, ()),                            # outer product
    ((5, 3), ()),
    ((128, 64), (48,)),
])
def test_direct_kernel_variants(case, accumulate, dtype):
    for seed in range(3):
        rng = np.random.default_rng(seed)
        got, want, kernel = _direct_case(rng, case[0], case[1], dtype, accumulate)
        tol = 1e-10 if dtype == np.complex128 else 1e-5
        scale = max(1.0, np.abs(want).max())
        assert np.abs(got - want).max() <= tol * scale, (case, kernel, seed)


@pytest.mark.parametrize("case", [
    ((64,), (64,), (8,), ()),
    ((2, 2, 2, 2, 2, 2, 2), (2, 2, 2, 2, 2, 2), (2, 2, 2, 2, 2), ()),
    ((7, 7, 3), (5, 7, 4), (7, 3, 2), (3,)),
    ((100,), (33,), (17,), (2, 2)),
    ((130,), (70,), (5,), ()),
    ((2,) * 10, (2,) * 9, (2,) * 7, ()),
    ((64,), (48,), (4096,), ()),                # few tiles, long K: split-K with FP64 atomics
    ((8, 8), (4, 4), (7, 7, 7, 7), (2, 2)),
])
def test_gemm_kernel_permuted_operands(case):
    rng = np.random.default_rng(3)
    got, want = _gemm_case(rng, *case)
    err = (got - want).abs().max().item()
    assert err <= 1e-10 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("case,kernel", [
    (((4, 4), (2, 2, 2, 2), (2,) * 13, (2,)), "gemm_thin"),        # 16 x 16, K = 8192, batch 2
    (((8,), (5,), (4, 4, 4, 4, 4, 3), ()), "gemm_thin"),             # T = 1, ragged K tail
    (((5, 5), (3, 7), (2, 3, 5, 7, 11), (3,)), "gemm_thin"),         # T = 4, ragged M / N / K
    (((4,) * 5, (2,) * 9, (4,) * 4, ()), "gemm_ws"),                 # 1024 x 512 x 256: 128 tiles
    (((9, 9), (70,), (150,), (2,)), "gemm_ws"),                      # ragged tiles and K tail
    (((64,), (48,), (4096,), ()), "gemm_ws_splitk"),                 # one tile, 256 chunks: split-K items
    (((5, 13), (3, 11), (4, 4, 4, 4, 4, 2), (2,)), "gemm_ws_splitk"),   # ragged 65 x 33, K = 2048, batch 2
    (((2,) * 8, (2,) * 5, (2,) * 6, (2,)), "gemm_ws32"),             # 256 x 32 x 64, batch 2
    (((7, 7, 3), (5, 7), (3, 7), (3,)), "gemm_ws32"),                # ragged rows / columns / K
    (((130,), (9,), (40,), ()), "gemm_ws16"),
    (((4, 4, 4, 4, 2), (4, 4, 4, 4), (4, 4, 2), ()), "gemm_ws"),     # 32 tiles of 2 chunks
    (((2,) * 14, (2,) * 4, (2,) * 4, (2, 2)), "gemm_ws16"),          # 16384 x 16 x 16, batch 4
    (((4,) * 7, (4, 4, 2), (4, 4, 4), (2,)), "gemm_ws32"),           # 16384 x 32 x 64, batch 2
    (((13, 11, 9), (5, 4), (5, 7), (2, 2, 2, 2)), "gemm_ws32"),      # ragged 1287 x 20 x 35, batch 16
    (((13, 11, 9), (3, 3), (5, 7), (2, 2, 2, 2)), "gemm_ws16"),      # ragged 1287 x 9 x 35
    (((4,) * 6, (4,) * 5, (4, 4, 2), ()), "gemm_ws"),              # 1024 tiles, 2 chunks each
    (((4,) * 6, (4, 4, 4, 4, 2), (4, 4, 4, 4, 4, 2), (2,)), "gemm_ws"),  # K = 2048: two-level K offsets
    (((13, 11, 9), (7, 11, 3), (5, 7), (2, 2, 2)), "gemm_ws"),     # ragged tiles, K tail, batch
    (((4, 4, 4, 4, 2), (4, 4, 4, 4, 2), (4, 4, 2), (2, 2, 2)), "gemm_ws"),   # batch 8, 64 tiles each
    (((4,) * 6, (4,) * 5, (3,) * 7, ()), "gemm3m"),                  # K prefix never a multiple of 16: not ws
])
def test_gemm_thin_and_3m_kernels(case, kernel):
    """The r02 FP64 kernels: thin split-K pairs (whole M x N in one small tile) and the 3M
    complex multiplication (C_im = P3 - P1 - P2) -- same 1e-10 bar as the 4-product kernel."""
    rng = np.random.default_rng(5)
    got, want = _gemm_case(rng, *case)
    if kernel is not None:
        assert rt.load().b2_last_kernel().decode() == kernel
    err = (got - want).abs().max().item()
    assert err <= 1e-10 * max(1.0, want.abs().max().item()), err


@pytest.mark.parametrize("case,kernel", [
    (((4,) * 6, (4,) * 5, (4, 4, 2), ()), "gemm_ws"),                # 32-byte load + store pairs
    (((13, 11, 9), (7, 11, 3), (5, 7), (2, 2, 2)), "gemm_ws"),       # ragged: pair and single stores
    (((64,), (48,), (4096,), ()), "gemm_ws_splitk"),                 # atomics onto the running sum
    (((4, 4), (2, 2, 2, 2), (2,) * 13, (2,)), "gemm_thin"),
    (((13, 11, 9), (3, 3), (5, 7), (2, 2, 2, 2)), "gemm_ws16"),
])
def test_gemm_kernels_accumulate_into_c(case, kernel):
    """accumulate = 1 (a slice or a formal-sum term added to the running result): C keeps its
    values and receives alpha * A.B on top -- no zero fill in the split-K kernels."""
    rng = np.random.default_rng(6)
    got, want = _gemm_case(rng, *case, accumulate=1)
    assert rt.load().b2_last_kernel().decode() == kernel
    err = (got - want).abs().max().item()
    assert err <= 1e-10 * max(1.0, want.abs().max().item()), err


@pytest.mark.parametrize("case", [
    ((128,), (64,), (16,), ()),                 # one tile, one chunk
    ((128,), (32,), (64,), ()),
    ((256,), (128,), (48,), ()),
    ((2, 2, 2, 2, 2, 2, 2), (2, 2, 2, 2, 2, 2), (2, 2, 2, 2, 2), ()),
    ((7, 7, 3), (5, 7, 4), (7, 3, 2), (3,)),
    ((100,), (33,), (17,), (2, 2)),
    ((130,), (70,), (5,), ()),
    ((2,) * 10, (2,) * 9, (2,) * 7, ()),
    ((300,), (200,), (1000,), ()),
])
def test_gemm_c64_tcgen05_permuted_operands(case):
    """complex64 tensor-core path (tcgen05 kind::tf32, 3-product split): FP32-level
    accuracy, far inside the 1e-5 bar of the north star."""
    rng = np.random.default_rng(4)
    got, want = _gemm_case(rng, *case, dtype=np.complex64)
    assert rt.load().b2_last_kernel().decode() == "gemm_c64_tcgen05"
    err = (got - want).abs().max().item()
    assert err <= 5e-6 * max(1.0, want.abs().max().item()), err


def test_probs_amp_and_dm_match_oracle():
    torch = _torch()
    lib = rt.load()
    rng = np.random.default_rng(5)
    stream = torch.cuda.current_stream().cuda_stream
    # AMP
    t = rng.standard_normal((3, 4, 5)) + 1j * rng.standard_normal((3, 4, 5))
    t[rng.random(t.shape) < 0.3] = 0
    dt = torch.from_numpy(t).cuda()
    p = torch.empty(t.size, dtype=torch.float64, device="cuda")
    nz = torch.empty(t.size, dtype=torch.uint8, device="cuda")
    rt.check(lib.b2_probs_amp(t.size, dt.data_ptr(), p.data_ptr(), nz.data_ptr(), rt.B2_C128, stream))
    want = oracle_eval.prob_dist_pure(t)
    got = {tuple(int(i) for i in np.unravel_index(k, t.shape)): float(p[k])
           for k in np.flatnonzero(nz.cpu().numpy())}
    assert got.keys() == want.keys()
    for k in want:
        assert abs(got[k] - want[k]) <= 1e-15 * want[k]     # hypot()**2 within a few ulp
    # DM: axes (mode, qubit ket, qubit bra, bit, qmode ket, qmode bra)
    types = ["mode", "qubit", "bit", "qmode"]
    shape = (3, 2, 2, 2, 4, 4)
    kinds = (1, 2, 3, 1, 2, 3)
    dm = rng.standard_normal(shape) * (rng.random(shape) < 0.5) + 0j
    dm[0, 0, 0, 1, 2, 2] = -1e-13           # clamp branch
    dm[2, :, :, 1, :, :] = 0
    dm[2, 0, 1, 1, 0, 0] = 0.25             # key seen only off-diagonal: kept with p = 0
    ddm = torch.from_numpy(dm).cuda()
    n_meas = 3 * 2
    p = torch.empty(n_meas, dtype=torch.float64, device="cuda")
    seen = torch.empty(n_meas, dtype=torch.uint8, device="cuda")
    mx = torch.empty(1, dtype=torch.float64, device="cuda")
    dims = (C.c_int32 * 6)(*shape)
    kk = (C.c_int32 * 6)(*kinds)
    rt.check(lib.b2_probs_dm(6, dims, kk, ddm.data_ptr(), p.data_ptr(), seen.data_ptr(),
                             mx.data_ptr(), rt.B2_C128, stream))
    want = oracle_eval.prob_dist_mixed(dm, types)
    got = {tuple(int(i) for i in np.unravel_index(k, (3, 2))): float(p[k])
           for k in np.flatnonzero(seen.cpu().numpy())}
    assert got.keys() == want.keys()
    for k in want:
        assert abs(got[k] - want[k]) <= 1e-15 * max(1.0, abs(want[k]))
    assert want[(2, 1)] == 0.0 and got[(2, 1)] == 0.0


def test_fp64_peak_microbench_runs():
    torch = _torch()
    lib = rt.load()
    sink = torch.zeros(1, dtype=torch.float64, device="cuda")
    flops = C.c_double(0)
    for kind in (0, 1, 2):
        rt.check(lib.b2_fp64_peak_kernel(kind, 64, sink.data_ptr(), C.byref(flops), 0))
        assert flops.value > 0
    torch.cuda.synchronize()


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("n,fill", [(1, 1.0), (31, 0.5), (100_003, 0.01), (1 << 20, 0.0),
                                    (1 << 20, 0.001), (70_001, 1.0)])
def test_compact_nonzero_matches_flatnonzero(dtype, n, fill):
    """b2_compact_nonzero == np.flatnonzero + the values (unordered list); -0.0 and
    exact zeros are skipped, NaN is kept; an overflowing list reports count > cap."""
    torch = _torch()
    lib = rt.load()
    rng = np.random.default_rng(n + int(fill * 1000))
    host = np.zeros(n, dtype=dtype)
    mask = rng.random(n) < fill
    host[mask] = (rng.standard_normal(mask.sum()) + 1j * rng.standard_normal(mask.sum())).astype(dtype)
    if n > 40:
        host[7] = complex(-0.0, 0.0)
        host[11] = complex(0.0, -0.0)
        host[13] = complex(0.0, 3.0)
        host[n - 1] = complex(float("nan"), 0.0)
    want = np.flatnonzero(host)           # NaN != 0 -> kept
    dev = torch.from_numpy(host).cuda()
    for cap in (max(1, want.size), max(1, want.size // 2)):
        idx = torch.full((cap,), -1, dtype=torch.int64, device="cuda")
        val = torch.zeros(cap, dtype=dev.dtype, device="cuda")
        cnt = torch.full((1,), 123, dtype=torch.int64, device="cuda")
        rt.check(lib.b2_compact_nonzero(n, dev.data_ptr(), rt.dtype_code(dtype), cap, idx.data_ptr(),
                                        val.data_ptr(), cnt.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream), "compact")
        count = int(cnt.item())
        if want.size <= cap:
            assert count == want.size
            got_idx = idx[:count].cpu().numpy()
            order = np.argsort(got_idx)
            assert np.array_equal(got_idx[order], want)
            got_val = val[:count].cpu().numpy()[order]
            assert np.array_equal(got_val, host[want], equal_nan=True)
        else:
            assert count > cap            # overflow: the caller must copy densely
