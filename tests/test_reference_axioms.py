"""The reference's own algebraic tests for the tensors on the eval path, restated:
ZW axioms and lemmas (test/test_zw.py:35-408), dual-rail / ZX correspondences
(test_zw.py:445-478), bit-controlled boxes (test/test_feed_forward.py:118-187),
classical arithmetic on digits (test/test_classical.py).  Each is an equality of
two diagrams' arrays, so it pins leaf arrays, daggers, swaps, spider merging and
the light-cone truncation dims together.

Every test runs twice: ``oracle`` = host compile + the numpy oracle (CPU,
``-m "not gpu"``), ``b200`` = the same networks through the CUDA kernels
(``-m gpu``)."""
import itertools
import math

import numpy as np
import pytest

from cases import oracle_eval, oracle_nested
from oracle.contract import contract_network
from optyx_b200 import classical, photonic
from optyx_b200.core import control, diagram, zw, zx
from optyx_b200.core.channel import Channel
from optyx_b200.core.diagram import DualRail, Mode, PhotonThresholdDetector, Scalar, Spider, Swap, mode
from optyx_b200.core.zw import Create, Id, Select, W, ZBox


def _oracle(d, input_dims=None):
    if hasattr(d, "terms"):
        return oracle_eval(d)
    return contract_network(d.to_tensor(input_dims, nested_eval=oracle_nested))


def _b200(d, input_dims=None):
    from optyx_b200 import executor
    from optyx_b200.core.backends import B200Backend
    backend = B200Backend()
    if hasattr(d, "terms"):
        nets = d.to_networks(nested_eval=backend._nested_eval)
        return sum(executor.evaluate(n).cpu().numpy() for n in nets)
    net = d.to_tensor(input_dims, nested_eval=backend._nested_eval)
    return executor.evaluate(net).cpu().numpy()


@pytest.fixture(params=["oracle", pytest.param("b200", marks=pytest.mark.gpu)])
def ev(request):
    return _oracle if request.param == "oracle" else _b200


def same(a, b, tol=1e-8):
    """``compare_arrays_of_different_sizes`` (utils.py:127-137): equal on the common
    flat prefix -- and, beyond the reference's check, equal on the common leading
    block when both have the same rank (the truncations of two diagrams that agree
    in infinite dimension differ only in how far each axis reaches)."""
    a, b = np.asarray(a), np.asarray(b)
    fa, fb = a.reshape(-1), b.reshape(-1)
    n = min(fa.size, fb.size)
    ok = not np.any(np.abs(fa[:n] - fb[:n]) > tol)
    if a.ndim == b.ndim and a.ndim > 0:
        blk = tuple(slice(0, min(x, y)) for x, y in zip(a.shape, b.shape))
        ok = ok and not np.any(np.abs(a[blk] - b[blk]) > tol)
    return ok


FS = [lambda i: i, lambda i: math.factorial(i), lambda i: np.exp(i)]
LEGS = list(itertools.product(range(len(FS)), range(1, 3), range(1, 3), range(1, 3), range(1, 3),
                              range(1, 3)))


@pytest.mark.parametrize("f,a_in,b_in,a_out,b_out,between", LEGS)
def test_spider_fusion(ev, f, a_in, b_in, a_out, b_out, between):
    fs = FS[f]
    left = ZBox(a_in, a_out + between, fs) @ Id(b_in) \
        >> Id(a_out) @ ZBox(b_in + between, b_out, fs)
    right = ZBox(a_in + b_in, a_out + b_out, lambda i: fs(i) * fs(i))
    # the reference's flat-prefix comparison only: with one input wire left at the
    # default dim the two sides legitimately truncate the outputs differently
    la, ra = ev(left).reshape(-1), ev(right).reshape(-1)
    n = min(la.size, ra.size)
    assert not np.any(np.abs(la[:n] - ra[:n]) > 1e-8)


@pytest.mark.parametrize("legs", [1, 2])
def test_z_conj(ev, legs):
    assert same(ev(ZBox(legs, legs, lambda i: 1j).dagger()), ev(ZBox(legs, legs, lambda i: -1j)))


def test_w_symmetry_and_associativity(ev):
    assert same(ev(W(2)), ev(W(2) >> Swap(mode, mode)))                       # bSym
    assert same(ev(W(2), [5]), ev(W(2) >> Swap(mode, mode), [5]))             # with input dims
    assert same(ev(W(2) >> W(2) @ Id(1)), ev(W(2) >> Id(1) @ W(2)))           # bAso


def test_swaps_identities_permutations(ev):
    assert same(ev(Swap(mode, mode) >> Swap(mode, mode)), ev(Id(2)))
    assert same(ev(Swap(mode, mode).dagger()), ev(Swap(mode, mode)))
    for k in (1, 2):
        assert same(ev(Id(k).dagger()), ev(Id(k)))
    perm = diagram.Diagram.permutation([1, 0], Mode(2))
    assert same(ev(perm), ev(perm.dagger()))


def test_bialgebra_and_unit(ev):
    bba_l = W(2) @ W(2) >> Id(1) @ Swap(mode, mode) @ Id(1) >> W(2).dagger() @ W(2).dagger()
    assert same(ev(bba_l), ev(W(2).dagger() >> W(2)))                         # bBa
    assert same(ev(W(2) >> Select(0) @ Id(1)), ev(Id(1)))                     # bId


def test_bzba(ev):
    n = [float(np.sqrt(math.factorial(i))) for i in range(5)]
    frac = [float(1 / np.sqrt(math.factorial(i))) for i in range(5)]
    left = (ZBox(1, 2, n) @ ZBox(1, 2, n) >> Id(1) @ Swap(mode, mode) @ Id(1)
            >> W(2).dagger() @ W(2).dagger() >> Id(1) @ ZBox(1, 1, frac))
    assert same(ev(left), ev(W(2).dagger() >> ZBox(1, 2, lambda i: 1)))
    left = (ZBox(1, 1, n) @ ZBox(1, 1, n) >> Spider(1, 2, Mode(1)) @ Spider(1, 2, Mode(1))
            >> Id(1) @ Swap(mode, mode) @ Id(1) >> W(2).dagger() @ W(2).dagger()
            >> Id(1) @ ZBox(1, 1, frac))
    assert same(ev(left), ev(W(2).dagger() >> Spider(1, 2, Mode(1))))


def test_copy_scalars_bone(ev):
    assert same(ev(Create(4) >> ZBox(1, 2, lambda i: 1)), ev(Create(4) @ Create(4)))    # K0_infty
    assert same(ev(Create(1) >> ZBox(1, 1, [1, 2]) >> Select(1)), ev(ZBox(0, 0, [2])))
    assert same(ev(Create(1) >> ZBox(1, 1, lambda i: i) >> Select(1)),
                ev(ZBox(0, 0, lambda i: i + 1)))
    assert same(ev(Create(1) >> Select(0)), [0])
    assert same([0], ev(Create(0) >> Select(1)))


def test_branching(ev):
    assert same(ev(Create(1) >> W(2)), ev(Create(1) @ Create(0) + Create(0) @ Create(1)))


@pytest.mark.parametrize("k", [1, 2, 3])
def test_normalisation_and_lemma_b6(ev, k):
    left = Create(k) @ ZBox(0, 0, [np.sqrt(math.factorial(k))])
    right = Id(0)
    for _ in range(k):
        right = right @ Create(1)
    assert same(ev(left), ev(right >> W(k).dagger()))
    assert same(ev(Create(k) >> ZBox(1, 1, lambda i: i + 1)), ev(Create(k) @ ZBox(0, 0, [k + 1])))


def test_lemma_b8_b7_prop_54(ev):
    b8_l = (Create(1) >> ZBox(1, 2, [1, 1]) >> W(2) @ W(2) >> Id(1) @ ZBox(2, 0, [1, 1]) @ Id(1))
    b8_r = Create(1) >> W(2) >> ZBox(1, 2, [1, 1]) @ ZBox(1, 0, [1, 1])
    assert same(ev(b8_l), ev(b8_r))
    one = lambda i: 1                                                          # noqa: E731
    b7_l = Id(1) @ W(2).dagger() >> ZBox(2, 0, one)
    b7_r = (W(2) @ Id(2) >> Id(1) @ Id(1) @ Swap(mode, mode) >> Id(1) @ Swap(mode, mode) @ Id(1)
            >> ZBox(2, 0, one) @ ZBox(2, 0, one))
    assert same(ev(b7_l), ev(b7_r))
    p_l = (Create(1) @ Id(1) >> ZBox(1, 2, one) @ Id(1) >> Id(1) @ W(2).dagger() >> Id(1) @ W(2)
           >> ZBox(2, 0, one) @ Id(1))
    p_r = (Create(1) @ Id(1) >> ZBox(1, 2, one) @ Id(1) >> Swap(mode, mode) @ Id(1)
           >> W(2) @ W(2) @ W(2) >> Id(1) @ ZBox(2, 0, one) @ ZBox(2, 0, one) @ Id(1)
           >> W(2).dagger())
    assert same(ev(p_l), ev(p_r))


# ---- test_zw.py:445-478: dual rail <-> ZX -------------------------------------------
def test_dual_rail_x_states(ev):
    assert np.allclose(ev(zx.X(0, 1) @ Scalar(0.5 ** 0.5) >> DualRail()), ev(Create(1, 0)))
    assert np.allclose(ev(zx.X(0, 1, phase=0.5) @ Scalar(0.5 ** 0.5) >> DualRail()),
                       ev(Create(0, 1)))


def test_dual_rail_beamsplitter_is_hadamard(ev):
    r = 0.5 ** 0.5
    bs = (W(2) @ W(2)
          >> ZBox(1, 1, lambda i: r ** i) @ ZBox(1, 1, lambda i: r ** i)
          @ ZBox(1, 1, lambda i: r ** i) @ ZBox(1, 1, lambda i: (-r) ** i)
          >> Id(1) @ zw.SWAP @ Id(1) >> W(2).dagger() @ W(2).dagger())
    left = ev(DualRail() >> bs)
    right = ev(zx.H >> DualRail())
    blk = tuple(slice(0, min(x, y)) for x, y in zip(left.shape, right.shape))
    assert np.allclose(left[blk], right[blk])
    rest = left.copy()
    rest[blk] = 0
    assert np.allclose(rest, 0)            # one photon in: nothing beyond occupation 1


@pytest.mark.parametrize("phase", [0.0, 0.3, 0.6])
def test_dual_rail_phase_shift(ev, phase):
    left = DualRail() >> Mode(1) @ photonic.Phase(phase).get_kraus()
    assert np.allclose(ev(zx.Z(1, 1, phase=phase) >> DualRail()), ev(left))


# ---- test_feed_forward.py:118-187: bit-controlled boxes --------------------------------
def _circuits():
    return [
        (photonic.Phase(0.1).get_kraus(), None),
        (photonic.Phase(0.456).get_kraus(), photonic.Phase(0.8765).get_kraus()),
        (photonic.BS.get_kraus(), None),
        (photonic.BS.get_kraus(), photonic.MZI(0.324, 0.9875).get_kraus()),
        (W(2).dagger() >> W(2), photonic.MZI(0.324, 0.9875).get_kraus()),
    ]


@pytest.mark.parametrize("which", range(5))
def test_binary_controlled_box(ev, which):
    action, default = _circuits()[which]
    box = control.BitControlledBox(action) if default is None else \
        control.BitControlledBox(action, default)
    on = (Create(1) >> PhotonThresholdDetector()) @ Mode(len(action.cod)) >> box
    off = (Create(0) >> PhotonThresholdDetector()) @ Mode(len(action.cod)) >> box
    assert np.allclose(ev(action), ev(on))
    if default is not None:
        assert np.allclose(ev(default), ev(off))


@pytest.mark.parametrize("which", range(5))
def test_bit_controlled_gate_doubled(ev, which):
    action, default = _circuits()[which]
    action = Channel("action", action)
    default = Channel("default", default) if default is not None else None
    gate = classical.BitControlledGate(action) if default is None else \
        classical.BitControlledGate(action, default)
    head = lambda n: (photonic.Create(n) >> photonic.PhotonThresholdMeasurement()) \
        @ photonic.Id(len(action.cod))                                         # noqa: E731
    assert np.allclose(ev(action.double()), ev((head(1) >> gate).double()))
    if default is not None:
        assert np.allclose(ev(default.double()), ev((head(0) >> gate).double()))


# ---- test_classical.py: arithmetic on digits, doubled -----------------------------------
@pytest.mark.parametrize("left,right", [
    (lambda: classical.Digit(2, 3) >> classical.AddN(2), lambda: classical.Digit(5)),
    (lambda: classical.Digit(2, 5) >> classical.SubN(), lambda: classical.Digit(3)),
    (lambda: classical.Digit(2, 3) >> classical.MultiplyN(), lambda: classical.Digit(6)),
    (lambda: classical.Digit(6, 2) >> classical.DivideN(), lambda: classical.Digit(3)),
    (lambda: classical.Digit(5) >> classical.Mod2(), lambda: classical.Bit(1)),
    (lambda: classical.Digit(2) >> classical.CopyN(3), lambda: classical.Digit(2, 2, 2)),
    (lambda: classical.Digit(2, 3) >> classical.SwapN(), lambda: classical.Digit(3, 2)),
    (lambda: classical.Bit(1, 0) >> classical.PostselectBit(1, 0), lambda: classical.Scalar(1)),
    (lambda: classical.Digit(2, 3, 4) >> classical.PostselectDigit(2, 3, 4),
     lambda: classical.Scalar(1)),
    (lambda: classical.Bit(1) >> classical.NotBit(), lambda: classical.Bit(0)),
    (lambda: classical.Bit(1, 0) >> classical.XorBit(), lambda: classical.Bit(1)),
])
def test_classical_arithmetic(ev, left, right):
    la, ra = ev(left().double()).reshape(-1), ev(right().double()).reshape(-1)
    n = min(la.size, ra.size)
    assert not np.any(np.abs(la[:n] - ra[:n]) > 1e-8)


# ---- test_zw_2_circuit.py: multi-photon inputs through BBS / TBS / MZI ---------------------
# (the reference compares with Perceval's SLOS; here the independent second algorithm is
# the oracle's permanent formula, path.py:345-384)
def _lo_cases():
    out = []
    for p1, p2 in ((1, 2), (2, 1)):
        out.append(("BBS0", p1, p2, lambda: photonic.BBS(0)))
        for bias in (0, 0.5):
            out.append((f"BBS{bias}", p1, p2, lambda bias=bias: photonic.BBS(bias)))
            out.append((f"TBS{bias}", p1, p2, lambda bias=bias: photonic.TBS(bias)))
    for p1, p2, theta, phi in itertools.product((1, 2), (1, 2), (0, 1, 0.5), (0, 1, 0.5)):
        out.append((f"MZI{theta},{phi}", p1, p2, lambda t=theta, p=phi: photonic.MZI(t, p)))
    return out


@pytest.mark.parametrize("name,p1,p2,gate", _lo_cases())
def test_two_mode_gates_with_multi_photon_inputs_match_permanents(ev, name, p1, p2, gate):
    from oracle.permanent import lo_amplitude
    g = gate()
    arr = ev(Create(p1, p2) >> g.get_kraus())
    U = np.asarray(g.array, dtype=complex).reshape(2, 2)
    n = p1 + p2
    for out in np.ndindex(n + 1, n + 1):
        want = lo_amplitude(U, (p1, p2), out) if sum(out) == n else 0.0
        inside = all(o < s for o, s in zip(out, arr.shape))
        got = arr[out] if inside else 0.0          # an axis the light cone cut shorter
        assert abs(got - want) <= 1e-10, (name, out, got, want, arr.shape)


# ---- test_zw_2_circuit.py:126-215: the same gates as channels, post-selected, doubled ---------
# (CPU only: scalar probabilities of the CP-doubled diagram against |permanent amplitude|^2)
def _channel_cases():
    out = []
    for p1, p2 in ((1, 2), (2, 1)):
        for bias in (0, 0.5):
            out.append((f"BBS{bias}", p1, p2, lambda bias=bias: photonic.BBS(bias)))
            out.append((f"TBS{bias}", p1, p2, lambda bias=bias: photonic.TBS(bias)))
    for p1, p2, theta, phi in itertools.product((1, 2), (1, 2), (0, 0.5), (0, 1, 0.5)):
        out.append((f"MZI{theta},{phi}", p1, p2, lambda t=theta, p=phi: photonic.MZI(t, p)))
    return out


@pytest.mark.parametrize("name,p1,p2,gate", _channel_cases())
def test_post_selected_channel_probability_matches_permanent(name, p1, p2, gate):
    from oracle.permanent import lo_amplitude
    g = gate()
    d = photonic.Create(p1, p2) >> g >> classical.Select(p1, p2)
    prob = _oracle(d.double())
    U = np.asarray(g.array, dtype=complex).reshape(2, 2)
    want = abs(lo_amplitude(U, (p1, p2), (p1, p2))) ** 2
    assert np.ndim(prob) == 0
    assert abs(complex(prob) - want) <= 1e-10, (name, prob, want)


# ---- test_feed_forward.py:190-330: classical function boxes, controlled phase shifts ----------
# (CPU only)
def _arr_and_ndom(d, input_dims=None):
    net = d.to_tensor(input_dims, nested_eval=oracle_nested)
    return contract_network(net), net.n_dom


def _tensor_dagger(arr, n_dom):
    """``tensor.dagger()`` of an array whose axes are (dom..., cod...)."""
    nd = arr.ndim
    return np.conj(arr).transpose(list(range(n_dom, nd)) + list(range(n_dom)))


def _close_on_overlap(a, b):
    if a.shape != b.shape:
        blk = tuple(slice(0, min(x, y)) for x, y in zip(a.shape, b.shape))
        return a.ndim == b.ndim and np.allclose(a[blk], b[blk])
    return np.allclose(a, b)


def _dagger_consistent(d):
    arr, n_dom = _arr_and_ndom(d)
    return _close_on_overlap(_tensor_dagger(arr, n_dom), _arr_and_ndom(d.dagger())[0])


CLASSICAL_FUNCTIONS = [
    (lambda x: [x[0] ^ x[1]], [1, 1]),
    (lambda x: [x[0], x[0]], [[1], [1]]),
    (lambda x: [x[0] ^ x[1], x[0]], [[1, 1], [1, 0]]),
    (lambda x: [x[0] ^ x[1], x[0] ^ x[1]], [[1, 1], [1, 1]]),
]


@pytest.mark.parametrize("which", range(4))
def test_logical_matrix_box(which):
    function, m = CLASSICAL_FUNCTIONS[which]
    m = np.array(m)
    if m.ndim == 1:
        m = m.reshape(1, -1)
    cod, dom = diagram.Bit(len(m)), diagram.Bit(len(m[0]))
    f_arr = _oracle(control.ClassicalFunctionBox(function, dom, cod))
    assert np.allclose(f_arr, _oracle(control.BinaryMatrixBox(m)))
    assert _dagger_consistent(control.BinaryMatrixBox(m))


REAL_FUNCS = [lambda x: [x[0] * 0.234, x[0] * 0.912, x[0] * 0.184], lambda x: [0.23 * x[0]],
              lambda x: [0.8362, 0.193, 0.654]]


@pytest.mark.parametrize("which", range(3))
def test_controlled_phase_shift_dagger_and_numeric(which):
    f = REAL_FUNCS[which]
    n = len(f([0]))
    assert _dagger_consistent(control.ControlledPhaseShift(f, n))
    assert _dagger_consistent(Mode(n) @ Create(4) >> control.ControlledPhaseShift(f, n))
    for x in range(5):
        diag = Create(x) @ Mode(n) >> control.ControlledPhaseShift(f, n)
        zbox = diagram.Id(Mode(0))
        for y in f([x]):
            zbox = zbox @ ZBox(1, 1, lambda i, y=y: np.exp(2 * np.pi * 1j * y) ** i)
        assert np.allclose(_oracle(zbox), _oracle(diag))
        assert np.allclose(_oracle(zbox.dagger()), _oracle(diag.dagger()))
        assert _dagger_consistent(zbox)


@pytest.mark.parametrize("f,n", [(lambda x: [x[0]], 1),
                                 (lambda x: [0.23 * x[0], 0.456 * x[0], 0.876 * x[0], 0.654 * x[0]], 4),
                                 (lambda x: [0.23 * x[0]], 1),
                                 (lambda x: [0.23 * x[0], 0.456 * x[0], 0.876 * x[0]], 3)])
def test_controlled_phase_shift_dagger_2(f, n):
    assert _dagger_consistent(control.ControlledPhaseShift(f, n))


def test_photon_threshold_detector_dagger():
    assert _dagger_consistent(PhotonThresholdDetector())


def test_matrix_to_zw():
    """test_feed_forward.py:321-337"""
    d1 = photonic.matrix_to_zw(np.array([[1, 1], [1, 1]]))
    d2 = (W(2) @ W(2) >> Mode(1) @ Swap(Mode(1), Mode(1)) @ Mode(1) >> W(2).dagger() @ W(2).dagger())
    assert np.allclose(_oracle(d1), _oracle(d2))
