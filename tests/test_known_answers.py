"""CPU: the host-side diagram compiler (double / to_tensor / light-cone dims /
box rules) + the numpy oracle reproduce every known answer the reference's own
tests and docstrings pin for the eval path.  This is what pins
the oracle; the ``-m gpu`` tests then hold the CUDA path to the oracle."""
import itertools

import numpy as np
import pytest

from cases import (cfg1_interferometer, cfg3_interferometer, compile_channel, distinguishable_bs,
                   fusion_i_spider, fusion_ii_teleportation, ghz, lossy_hom, oracle_eval,
                   oracle_nested, random_clifford_t, random_internal_state, statevector_clifford_t)
from oracle import evalresult as ev
from oracle import leaves as oracle_leaves
from oracle.contract import contract_network
from oracle.permanent import lo_amplitude
from optyx_b200 import classical, photonic, qubits
from optyx_b200.core import channel, control, diagram, zw, zx
from optyx_b200.core.channel import Channel, CQMap, Discard, Measure, bit, mode, qmode, qubit
from optyx_b200.core.diagram import Id, Mode, Swap


def dbl(d):
    return contract_network(d.double().to_network(nested_eval=oracle_nested))


# ---- test/test_zw.py:410-442 ------------------------------------------------------
@pytest.mark.parametrize("s1,s2,amp", [(1, 1, 0), (2, 0, np.sqrt(2) * 1j / 2),
                                       (0, 2, np.sqrt(2) * 1j / 2)])
def test_hong_ou_mandel_amplitudes(s1, s2, amp):
    zb_i = zw.ZBox(1, 1, lambda i: (np.sin(0.25 * np.pi) * 1j) ** i)
    zb_1 = zw.ZBox(1, 1, lambda i: (np.cos(0.25 * np.pi)) ** i)
    bs = (zw.W(2) @ zw.W(2) >> zb_i @ zb_1 @ zb_1 @ zb_i
          >> Id(Mode(1)) @ Swap(diagram.mode, diagram.mode) @ Id(Mode(1))
          >> zw.W(2).dagger() @ zw.W(2).dagger())
    hom = zw.Create(1) @ zw.Create(1) >> bs >> zw.Select(s1) @ zw.Select(s2)
    assert np.allclose(oracle_eval(hom), amp, atol=1e-8)


# ---- README.md:176-178, photonic.py:131-138, 165-177 ----------------------------
def test_lossy_hom_probabilities():
    arr = oracle_eval(lossy_hom(0.8))
    p = ev.prob_dist(arr, ["mode"], ev.DM)
    assert p[(1,)] == pytest.approx(0.2, abs=1e-12)
    assert p[(2,)] == pytest.approx(0.8, abs=1e-12)


def test_backends_docstring_bs():
    # backends.py:41-47: single_prob((2, 0)) == 0.5
    arr = oracle_eval(photonic.Create(1, 1) >> photonic.BS)
    p = ev.prob_dist(arr, ["qmode", "qmode"], ev.AMP)
    assert round(p[(2, 0)], 1) == 0.5 and round(p[(0, 2)], 1) == 0.5 and (1, 1) not in p or \
        abs(p.get((1, 1), 0)) < 1e-12


# ---- photonic.py:1014-1037 ----------------------------------------------------------
def test_photon_loss_composition_and_limits():
    assert np.allclose(dbl(photonic.PhotonLoss(0.25)),
                       dbl(photonic.PhotonLoss(0.5) >> photonic.PhotonLoss(0.5)))
    assert np.allclose(dbl(photonic.Create(1) >> photonic.PhotonLoss(0.0)),
                       dbl(photonic.Create(0)))
    assert np.allclose(dbl(photonic.Create(1) >> photonic.PhotonLoss(1.0)),
                       dbl(photonic.Create(1)))


# ---- test/test_conjugation.py ---------------------------------------------------------
@pytest.mark.parametrize("gate", [photonic.Phase(0.27), photonic.BBS(0.27), photonic.TBS(0.27),
                                  photonic.MZI(0.27, 0.76)])
def test_causality_lo_gates(gate):
    dom = gate.dom
    assert np.allclose(dbl(Discard(dom)), dbl(gate >> Discard(dom)))


@pytest.mark.parametrize("box", [zx.Z, zx.X])
def test_causality_zx(box):
    gate = box(1, 1, 0.69)
    assert np.allclose(dbl(Discard(qubit)), dbl(Channel("Gate", gate) >> Discard(qubit)))


def test_causality_zw_kraus():
    d = photonic.TBS(0.24, 0.86).get_kraus()
    assert np.allclose(dbl(Discard(qmode ** 2)), dbl(Channel("diagram", d) >> Discard(qmode ** 2)))


# ---- test/test_qubit.py:102-115 (pins the zx.ZBox conjugation quirk) ----------------
def test_discard_qubits():
    assert np.allclose(dbl(qubits.Discard(2)), dbl(qubits.Z(1, 1) ** 2 >> qubits.Discard(2)))


def test_bit_flip_and_dephasing_goldens():
    a = oracle_eval(qubits.Z(0, 1) >> qubits.BitFlipError(0.43))
    assert np.allclose(a, np.array([[-.14, -0.14], [-0.14, -0.14]]))
    a = oracle_eval(qubits.Z(0, 1) >> qubits.DephasingError(0.43))
    assert np.allclose(a, np.array([[-.14, 1.], [1, -0.14]]))


def test_ket_bra():
    a = oracle_eval(qubits.Ket(1).get_kraus())
    b = oracle_eval(zx.X(0, 1, 0.5) @ diagram.Scalar(1 / np.sqrt(2)))
    assert np.allclose(a, b) and np.allclose(a, [0, 1])
    assert np.allclose(oracle_eval(qubits.Bra(0).get_kraus()), [1, 0])


# ---- README.md:50-93 teleportation, test/test_fusion.py ---------------------------------
def test_teleportation_is_identity():
    bell = qubits.Scalar(0.5 ** 0.5) @ qubits.Z(0, 2)
    for corr_x, corr_z in [(classical.CtrlX, classical.CtrlZ),
                           (classical.BitControlledGate(qubits.X(1, 1, 0.5)),
                            classical.BitControlledGate(qubits.Z(1, 1, 0.5)))]:
        tele = (qubit @ bell >> qubits.gate("CX") @ qubit >> qubits.H() @ qubit ** 2
                >> qubits.Measure(1) @ qubits.Measure(1) @ qubit >> bit @ corr_x >> corr_z)
        assert np.allclose(dbl(tele), dbl(qubits.Id(1)))


def test_fusion_ii_teleportation_and_fusion_i_spider():
    assert np.allclose(dbl(fusion_ii_teleportation()),
                       dbl(qubits.Id(1) @ qubits.Scalar(0.5 ** 0.5)))
    assert np.allclose(dbl(fusion_i_spider()), dbl(qubits.Z(2, 1)))


# ---- test/test_distinguishability.py:47-94 (seed 2025) -----------------------------------
def _internal_states():
    rng = np.random.default_rng(seed=2025)
    fixed = [np.array([1, 1]) / np.sqrt(2), np.array([1]), random_internal_state(rng)]
    pool = [random_internal_state(rng) for _ in range(5)]
    return fixed, pool


@pytest.mark.parametrize("k", range(3))
def test_indistinguishable_hom(k):
    s = _internal_states()[0][k]
    probs = dbl(distinguishable_bs(s, s))
    nz = {idx: val for idx, val in np.ndenumerate(np.round(probs, 6)) if val != 0}
    assert nz == {(0, 2): 0.5, (2, 0): 0.5}


@pytest.mark.parametrize("i,j", list(itertools.product(range(5), range(5)))[::3])
def test_partially_distinguishable_hom(i, j):
    pool = _internal_states()[1]
    s1, s2 = pool[i], pool[j]
    probs = np.round(dbl(distinguishable_bs(s1, s2)), 6)
    overlap = np.abs(s1 @ s2.conj()) ** 2
    assert np.isclose(probs[1, 1].real, 0.5 - 0.5 * overlap, atol=1e-4)


def test_dualrail_identity_and_threshold_detector():
    for state in [zx.Z(0, 1) @ diagram.Scalar(1 / np.sqrt(2)),
                  zx.X(0, 1, 0.3) @ diagram.Scalar(1 / np.sqrt(2))]:
        for internal in _internal_states()[0]:
            dr = diagram.DualRail(internal_state=internal).inflate(len(internal))
            got = oracle_eval(state >> dr >> dr.dagger())
            assert np.allclose(got, oracle_eval(state))
    d1 = (zw.Create(1, internal_states=([1, 0])) >> diagram.PhotonThresholdDetector()).inflate(2)
    d2 = zx.X(1, 0, 0.5) @ diagram.Scalar(0.5 ** 0.5)
    assert np.allclose(oracle_eval(d1), oracle_eval(d2))


# ---- test/test_backends.py:482-509: permanent cross-check for pure LO ---------------------
@pytest.mark.parametrize("name", ["BS", "phase_tbs", "MZI", "chip"])
def test_pure_lo_matches_permanent(name):
    circ = {"BS": photonic.BS,
            "phase_tbs": photonic.Phase(0.2) @ photonic.Phase(0.3) >> photonic.TBS(0.3),
            "MZI": photonic.MZI(0.2, 0.8),
            "chip": photonic.seeded_ansatz(4, 4, seed=7)}[name]
    n = len(circ.dom)
    d = photonic.Create(*(1,) * n) >> circ
    arr = oracle_eval(d)
    # single-particle matrix of the circuit from its gates
    U = np.eye(n, dtype=complex)
    for left, box, right in circ.layers():
        G = np.eye(n, dtype=complex)
        k = len(left)
        G[k:k + len(box.dom), k:k + len(box.dom)] = box.array.reshape(len(box.dom), len(box.cod))
        U = U @ G
    for out in np.ndindex(*arr.shape):
        want = lo_amplitude(U, (1,) * n, out) if sum(out) == n else 0.0
        assert abs(arr[out] - want) <= 1e-12, (out, arr[out], want)


# ---- test/test_zw_truncated_arrays.py:30-143: alternative W builder (arXiv:2306.02114) ----
@pytest.mark.parametrize("dims", [(3, 3, 3), (5, 5, 5), (7, 7, 7, 7), (9, 3, 5)])
def test_w_array_matches_alternative_builder(dims):
    from oracle.leaves import multinomial, occupation_numbers
    got = oracle_leaves.build_leaf(oracle_leaves.K_SUM, dims)
    want = np.zeros(dims, dtype=complex)
    for n in range(dims[0]):
        for config in occupation_numbers(n, len(dims) - 1):
            if all(c < d for c, d in zip(config, dims[1:])):
                want[(n,) + tuple(config)] += np.sqrt(multinomial(config))
    assert np.allclose(got, want)


# ---- test/test_classical.py style truth tables ----------------------------------------------
def test_classical_truth_tables():
    assert np.allclose(oracle_eval(classical.Digit(2, 3) >> classical.AddN(2)
                                   >> classical.PostselectDigit(5)), 1.0)
    assert np.allclose(oracle_eval(classical.Digit(3, 2) >> classical.MultiplyN()
                                   >> classical.PostselectDigit(6)), 1.0)
    assert np.allclose(oracle_eval(classical.Digit(9, 3) >> classical.DivideN()
                                   >> classical.PostselectDigit(3)), 1.0)
    for a in (0, 1):
        for b in (0, 1):
            st = classical.Bit(a, b)
            norm = oracle_eval(st >> classical.PostselectBit(a, b))
            for box, f in [(classical.AndBit(), a & b), (classical.OrBit(), a | b),
                           (classical.XorBit(), a ^ b)]:
                got = oracle_eval(st >> box >> classical.PostselectBit(f)) / norm
                bad = oracle_eval(st >> box >> classical.PostselectBit(1 - f)) / norm
                assert np.isclose(abs(got), 1.0 / np.sqrt(2) ** 0 * abs(got)) and abs(bad) < 1e-12
                assert abs(got) > 1e-6
    m = oracle_eval(control.BinaryMatrixBox([[1, 1]]))
    x = oracle_eval(zx.X(2, 1) @ diagram.Scalar(np.sqrt(2)))
    assert np.allclose(m, x)                      # control.py:384-393


def test_controlled_phase_shift_docstring():
    # control.py:225-238
    f = lambda x: [x[0] * 0.1, x[0] * 0.2, x[0] * 0.3]
    n = 3
    cps = zw.Create(2) @ Mode(n) >> control.ControlledPhaseShift(f, n_modes=n)
    zbox = Id(Mode(0))
    for y in f([2]):
        zbox = zbox @ zw.ZBox(1, 1, lambda i, y=y: np.exp(2 * np.pi * 1j * y) ** i)
    assert np.allclose(oracle_eval(cps), oracle_eval(zbox))


def test_bit_controlled_box_docstring():
    # control.py:74-88
    action = photonic.Phase(0.1).get_kraus()
    default = zw.ZBox(1, 1, lambda x: 1)
    a_test = (zw.Create(1) >> diagram.PhotonThresholdDetector()) @ Mode(1) \
        >> control.BitControlledBox(action)
    d_test = (zw.Create(0) >> diagram.PhotonThresholdDetector()) @ Mode(1) \
        >> control.BitControlledBox(default)
    assert np.allclose(oracle_eval(action), oracle_eval(a_test))
    assert np.allclose(oracle_eval(default), oracle_eval(d_test))


# ---- test/test_channel.py:22-34: noisy Bell fidelity (array box, CQMap) -------------------
def test_noisy_bell_fidelity():
    re = np.array([[0.012, 0.014, 0.014, 0.000], [0.014, 0.508, 0.475, 0.008],
                   [0.014, 0.475, 0.479, 0.009], [0.000, 0.008, 0.009, 0.000]])
    im = np.exp(1j * np.pi * np.array([[0.000, -1.850, -1.825, -0.985],
                                       [1.850, 0.000, -0.002, -0.902],
                                       [1.825, 0.002, 0.000, -0.931],
                                       [0.985, 0.902, 0.931, 0.000]]))
    bell_density = np.multiply(re, im)
    D, b = diagram.Diagram, diagram.bit
    bell = diagram.Box(name="Bell", dom=b ** 2, cod=b ** 2, array=bell_density)
    bell = (diagram.Spider(0, 2, b) >> D.id(b) @ diagram.Spider(0, 2, b) @ D.id(b)
            >> D.permutation([0, 1, 3, 2], b ** 4) >> D.id(b ** 2) @ bell
            >> D.permutation([0, 2, 1, 3], b ** 4))
    noisy = CQMap("Physical Bell", bell @ diagram.Scalar(1 / 0.999), dom=channel.Ty(),
                  cod=qubit ** 2)
    X = Channel("X", zx.X(1, 1, 0.5))
    effect = Channel("Perfect Bell Effect",
                     diagram.Spider(2, 0, b) @ diagram.Scalar(1 / np.sqrt(2)))
    fid = dbl(noisy >> channel.Diagram.id(qubit) @ X >> effect).real
    assert np.isclose(fid, .96898, rtol=1e-3)


# ---- BASELINE configs at CPU-checkable sizes ------------------------------------------------
def test_cfg1_probabilities_sum_to_one_and_photon_number():
    d = cfg1_interferometer(4, (1, 1, 0, 0), loss=0.9)
    arr = oracle_eval(d)
    p = ev.prob_dist(arr, ["mode"] * 4, ev.DM)
    assert sum(p.values()) == pytest.approx(1.0, abs=1e-10)
    assert all(sum(k) <= 2 for k, v in p.items() if v > 1e-14)
    # survival of both photons: 0.9 ** 2
    assert sum(v for k, v in p.items() if sum(k) == 2) == pytest.approx(0.81, abs=1e-10)


def test_cfg2_ghz_known_answer():
    n = 6
    arr = oracle_eval(ghz(n, measure=True))
    p = ev.prob_dist(arr, ["bit"] * n, ev.DM)
    assert p[(0,) * n] == pytest.approx(0.5) and p[(1,) * n] == pytest.approx(0.5)
    # with the reference's DephasingError quirk the raw diagonal carries (2p-1)^n,
    # prob_dist renormalises
    arr = oracle_eval(ghz(4, noise=0.95, measure=True))
    p = ev.prob_dist(arr, ["bit"] * 4, ev.DM)
    assert p[(0,) * 4] == pytest.approx(0.5) and p[(1,) * 4] == pytest.approx(0.5)


@pytest.mark.parametrize("n,depth", [(4, 3), (6, 4), (7, 5)])
def test_cfg4_amplitude_matches_statevector(n, depth):
    amp = oracle_eval(random_clifford_t(n, depth, seed=1))
    want = statevector_clifford_t(n, depth, seed=1)
    assert abs(amp - want) <= 1e-12


def test_cfg3_small_distribution_is_normalised():
    d = cfg3_interferometer(width=3, n_photons=2, measured=2, depth=2)
    arr = oracle_eval(d)
    p = ev.prob_dist(arr, ["mode", "mode"], ev.DM)
    assert sum(p.values()) == pytest.approx(1.0, abs=1e-9)
    assert all(v >= -1e-12 for v in p.values())


# ---- README.md:230-296 (cfg5a) and test_fusion.py + input phases (cfg5b) --------------------
def test_cfg5_fusion_link_fidelity_and_feed_forward_teleport():
    import math
    from cases import fusion_link, fusion_teleport_input
    ratios = []
    for i in (0, 7, 14):
        th = i * (math.pi / 2) / 14
        v = (math.cos(th), math.sin(th))
        ratios.append((oracle_eval(fusion_link(v, "bell"))
                       / oracle_eval(fusion_link(v, "discard"))).real)
    assert ratios[0] == pytest.approx(1.0, abs=1e-12)          # indistinguishable: perfect link
    assert ratios[0] > ratios[1] > ratios[2]                   # monotone in the overlap
    assert ratios[2] == pytest.approx(0.5, abs=1e-12)
    got = oracle_eval(fusion_teleport_input(0.3))
    want = dbl(qubits.Ket("+") >> qubits.Z(1, 1, 0.3) @ qubits.Scalar(0.5 ** 0.5))
    assert np.allclose(got, want)


def test_branching_sum_equals_w():
    """test_zw.py:284-292: Create(1) >> W(2) == Create(1) @ Create(0) + Create(0) @ Create(1)
    (formal sums: diagram.py:710-765, terms padded to a common shape and added)."""
    from optyx_b200.core import zw
    lhs = oracle_eval(zw.Create(1) >> zw.W(2))
    rhs_diagram = zw.Create(1) @ zw.Create(0) + zw.Create(0) @ zw.Create(1)
    assert len(rhs_diagram.terms) == 2
    rhs = oracle_eval(rhs_diagram)
    shape = tuple(max(a, b) for a, b in zip(lhs.shape, rhs.shape))
    pad = lambda a: np.pad(a, [(0, s - d) for s, d in zip(shape, a.shape)])      # noqa: E731
    assert np.allclose(pad(lhs), pad(rhs))
    # the product's own padding (EmbeddingTensor leaves) gives the same tensor
    nets = rhs_diagram.to_networks()
    from oracle.contract import contract_network
    assert np.allclose(sum(contract_network(n) for n in nets), rhs)


def test_sum_algebra_distributes_and_checks_types():
    from optyx_b200.core import zw
    a, b, c = zw.Create(1), zw.Create(2), zw.W(2)
    s = (a + b) >> c
    assert [len(t.boxes) for t in s.terms] == [2, 2] and s.cod == c.cod
    assert len(sum([a, b, a]).terms) == 3
    t = zw.Create(0) @ (a + b)
    assert len(t.terms) == 2 and len(t.cod) == 2
    with pytest.raises(ValueError):
        a + c


def test_bosonic_operator_hamiltonian_sum():
    """test_channel.py:36-64: a hopping Hamiltonian a_0^dag a_1 + a_1^dag a_0 as a
    formal sum of channels; its Kraus tensor and its CP-doubled evaluation describe
    the same operator (the doubled tensor is kraus (x) conj(kraus) summed per term)."""
    from optyx_b200.core.channel import Diagram
    matrix = [[0, 1], [1, 0]]
    terms = [Diagram.from_bosonic_operator(2, [(i, False), (j, True)], scalar=matrix[i][j])
             for i in range(2) for j in range(2)]
    ham = Diagram.sum_factory(terms)
    assert len(ham.terms) == 4 and ham.is_pure
    kraus = oracle_eval(ham)                       # pure: AMP tensor of the Kraus sum
    # a_0^dag a_1 |0,1> = |1,0>  and  a_1^dag a_0 |1,0> = |0,1>
    assert kraus.shape[:2] == (2, 2)
    assert kraus[0, 1, 1, 0] == pytest.approx(1.0) and kraus[1, 0, 0, 1] == pytest.approx(1.0)
    assert kraus[0, 0].sum() == 0 and kraus[1, 1, 1, 1] == pytest.approx(0.0)
    with pytest.raises(ValueError):
        Diagram.from_bosonic_operator(2, [(2, False)])


# ---- host compile internals: the one-pass functor equals the layer-by-layer definition -------
def _functor_by_layers(d, on_ob, on_box, factory):
    """F(d) as the reference builds it: id >> (F(left) @ F(box) @ F(right)) per layer."""
    res = factory.id(on_ob(d.dom))
    for left, box, right in d.layers():
        res = res >> (factory.id(on_ob(left)) @ on_box(box) @ factory.id(on_ob(right)))
    return res


@pytest.mark.parametrize("name", ["lossy_hom", "cfg1_4mode", "fusion_ii", "hom_distinguishable",
                                  "ghz5_noisy_open3"])
def test_one_pass_functors_equal_layer_by_layer_application(name):
    from cases import EVAL_CASES
    d = EVAL_CASES[name]()
    fast = d.double()
    slow = _functor_by_layers(d, lambda ty: ty.double(), lambda box: box.double(), diagram.Diagram)
    assert fast.dom == slow.dom and fast.cod == slow.cod
    assert fast.offsets == slow.offsets
    assert [type(b) for b in fast.boxes] == [type(b) for b in slow.boxes]
    assert np.allclose(contract_network(fast.to_network(nested_eval=oracle_nested)),
                       contract_network(slow.to_network(nested_eval=oracle_nested)))
    base = cfg3_interferometer(width=3, n_photons=2, measured=2, depth=2, inflate=False)
    fast = base.inflate(2)
    slow = _functor_by_layers(base, lambda ty: ty.inflate(2), lambda box: box.inflate(2),
                              channel.Diagram)
    assert fast.dom == slow.dom and fast.cod == slow.cod and fast.offsets == slow.offsets


# ---- optyx/photonic.py module docstring (lines 131-138, 165-177, 199-243) ---------------------
def test_photonic_module_docstring_examples():
    hom = photonic.Create(1, 1) >> photonic.BS >> classical.Select(1, 1)
    assert float(np.round(dbl(hom).real, 1)) == 0.0               # impossible for an ideal BS
    lossy = (photonic.Create(1, 1) >> photonic.PhotonLoss(0.2) @ photonic.Id(1) >> photonic.BS
             >> classical.Select(1, 0))
    assert float(np.round(dbl(lossy).real, 1)) == 0.4
    rng = np.random.default_rng(11)
    s1 = rng.random(2) + 1j * rng.random(2)
    s1 /= np.linalg.norm(s1)
    s2 = rng.random(2) + 1j * rng.random(2)
    s2 /= np.linalg.norm(s2)
    chan = (photonic.Create(1, 1, internal_states=(s1, s2)) >> photonic.BS
            >> photonic.NumberResolvingMeasurement(2)).inflate(2)
    result = dbl(chan)
    assert np.isclose(result[1, 1].real, 0.5 - 0.5 * abs(np.dot(s1, s2.conjugate())) ** 2)


# ---- gate docstrings (photonic.py:414-420, 556-570, 632-637, 726-741): g >> g.dagger() = id ----
@pytest.mark.parametrize("gate", [photonic.BBS(0.4), photonic.TBS(0.25), photonic.MZI(0.3, 0.7),
                                  photonic.Phase(0.35) @ photonic.Id(1), photonic.HadamardBS()])
@pytest.mark.parametrize("photons", [(1, 1), (2, 0), (1, 0), (0, 2)])
def test_gate_followed_by_its_dagger_is_identity(gate, photons):
    got = oracle_eval(photonic.Create(*photons) >> gate >> gate.dagger())
    want = oracle_eval(photonic.Create(*photons))
    n = sum(photons)
    for out in np.ndindex(*got.shape):
        ref = want[out] if all(o < s for o, s in zip(out, want.shape)) else 0.0
        assert abs(got[out] - ref) <= 1e-10, (photons, out, got[out], ref)
    assert abs(got[photons] - want[photons]) <= 1e-10 and abs(want[photons]) > 0 and n >= 1


# ---- optyx/classical.py module docstring (lines 93-131) --------------------------------------
def test_classical_module_docstring_examples():
    target = dbl(classical.XorBit(2))
    f = classical.ClassicalFunction(lambda b: [b[0] ^ b[1]], bit ** 2, bit)
    m = classical.BinaryMatrix([[1, 1]])
    assert np.allclose(dbl(f), target)
    assert np.allclose(dbl(m), target)
    parity = classical.AddN(2) >> classical.Mod2()
    assert np.allclose(dbl(classical.Digit(3, 2) >> parity >> classical.PostselectBit(1)), 1)


# ---- optyx/qubits.py:152-173: the docstring circuit, evaluated by hand ----
def test_qubits_docstring_teleportation_variant_by_hand():
    """Kraus operators by hand: K_rs = (-1)^{rs} Z^r X^s |r><r|, so the channel is
    rho -> Tr(rho) * 1 and the doubled tensor is delta(ket_in, bra_in) delta(ket_out, bra_out)."""
    from cases import teleportation
    got = oracle_eval(teleportation())
    want = np.einsum("ab,cd->abcd", np.eye(2), np.eye(2))
    assert got.shape == (2, 2, 2, 2) and np.allclose(got, want)
    # the identity channel is the same pair of Kronecker deltas under another axis order
    identity = np.einsum("ac,bd->abcd", np.eye(2), np.eye(2))
    assert not np.allclose(got, identity) and np.allclose(got.transpose(0, 2, 1, 3), identity)
