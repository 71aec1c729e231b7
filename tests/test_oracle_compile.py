"""The independent host-compile oracle (``oracle/compile.py``) against (a) the
reference-derived truncation dims, (b) the product's host compile -- box by box for
the truncation bookkeeping (exact) and end to end for the values.

``oracle/compile.py`` restates channel.py:277-358, 600-641 (purity, Kraus map, CP
doubling), diagram.py:277-284 (per-box conjugation), diagram.py:301-375 +
utils.py:347-510 (``to_tensor`` with the literal BACKWARD light-cone walk) and every
box's ``truncation`` without importing ``optyx_b200``; every known answer of the
reference reproduces through it (``cases.oracle_eval`` is built on it, so the whole of
tests/test_known_answers.py, test_reference_axioms.py and test_golden.py pin it).
"""
import math

import numpy as np
import pytest

from cases import (EVAL_CASES, cfg1_interferometer, cfg3_interferometer, compile_channel,
                   distinguishable_bs, fusion_i_spider, fusion_ii_teleportation, fusion_link,
                   fusion_teleport_input, ghz, lossy_hom, oracle_eval, product_compile_oracle_eval,
                   random_clifford_t, teleportation)
from oracle import compile as oc
from optyx_b200 import classical, photonic, qubits
from optyx_b200.core import channel, diagram, zw, zx
from optyx_b200.core.channel import qmode


def _more_cases():
    return {
        "teleportation": teleportation,
        "fusion_link_discard": lambda: fusion_link((math.cos(0.3), math.sin(0.3)), "discard"),
        "fusion_link_bell": lambda: fusion_link((math.cos(0.7), math.sin(0.7)), "bell"),
        "fusion_teleport_input": lambda: fusion_teleport_input(0.2),
        "numop": lambda: photonic.Create(2, 1) >> photonic.BS >> photonic.NumOp() @ photonic.Id(1),
        "threshold": lambda: (photonic.Create(1, 1, 1) >> photonic.seeded_ansatz(3, 3, 1)
                              >> photonic.PhotonThresholdMeasurement(1) @ photonic.Id(2)),
        "ghz5_noisy_open2": lambda: ghz(5, noise=0.9, open_qubits=2),
        "cfg1_5mode": lambda: cfg1_interferometer(5, (1, 1, 1, 0, 0)),
        "clifford_t_6x5_open": lambda: random_clifford_t(6, 5, seed=1, close=False),
        "hom_dist_complex": lambda: distinguishable_bs(np.array([0.6, 0.8j]), np.array([0.8, 0.6])),
    }


ALL = dict(EVAL_CASES, **_more_cases())


@pytest.mark.parametrize("name", sorted(ALL))
def test_product_host_compile_matches_independent_oracle(name):
    """The product's double() / to_tensor() / leaf descriptors (contracted on the CPU)
    against the independent restatement: same shape, values to 1e-12."""
    d = ALL[name]()
    want = oracle_eval(d)
    got = product_compile_oracle_eval(d)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max())


def _product_trace(d):
    """Per non-swap box of the tensor diagram the product compiles: (kind, dims_in, dims_out)."""
    from cases import oracle_nested
    td = d.get_kraus() if d.is_pure else d.double()
    trace = []
    diagram.compile_diagram(td, nested_eval=oracle_nested, trace=trace)
    return [(oc.core_rec(box).kind, di, do) for box, di, do in trace
            if not isinstance(box, diagram.Swap)]


@pytest.mark.parametrize("name", sorted(ALL))
def test_truncation_dims_match_literal_backward_walk(name):
    """SURVEY 8(a) row a3': the product's forward light-cone propagation gives, for
    every box, exactly the dims the literal backward walk of utils.py:436-510 gives
    (swaps are compared through the boxes after them: the two compilers order the
    elementary swaps of a permutation differently)."""
    d = ALL[name]()
    want = [t for t in oc.box_dims(d) if t[0] != "Swap"]
    got = _product_trace(d)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g == w, (g, w)


def _dims(d):
    return [(k, di, do) for k, di, do in oc.box_dims(d) if k != "Swap"]


def test_reference_derived_dims_lo():
    """Dims derived BY HAND from utils.py:436-510 + the boxes' determine_output_dimensions.

    ``Create(1, 1) >> BS`` (pure): Create -> [2, 2] (zw.py:462-470).  BS = matrix_to_zw
    (utils.py:33-87): W(2) on each input: the cone of wire 0 reaches Create output 0
    only: 1 photon + 1 = 2 -> outputs [2, 2]; Endo is uncapped (446-450); each W(2)^dagger
    sees one leg of each W: both Create outputs: 2 + 1 = 3, and sum(d - 1) + 1 = 3
    (zw.py:246-251) -> [3]."""
    got = _dims(photonic.Create(1, 1) >> photonic.BS)
    assert got[0] == ("Create", [], [2, 2])
    ws = [t for t in got if t[0] == "W"]
    assert ws == [("W", [2], [2, 2]), ("W", [2], [2, 2]), ("W", [2, 2], [3]), ("W", [2, 2], [3])]
    assert all(do == di for k, di, do in got if k == "Endo")
    # "truncation dim 4" of BASELINE cfg1 = 3 photons + 1 (utils.py:508-510)
    last = _dims(photonic.Create(1, 1, 1, 0, 0, 0) >> photonic.seeded_ansatz(6, 6, 0))
    assert max(max(do, default=0) for _, _, do in last) == 4
    assert oracle_eval(photonic.Create(1, 1, 1, 0, 0, 0) >> photonic.seeded_ansatz(6, 6, 0)).shape == (4,) * 6


def test_reference_derived_dims_open_inputs_and_non_lo():
    """An open mode input counts ``2 * input_dim`` (utils.py:508): W(2) on a default
    (dim 2) input has bound 2 * 2 + 1 = 5, its outputs keep dim 2 (zw.py:246-251).
    ``Create(3) >> W(2)``: 3 + 1 = 4 -> [4, 4].  A NON_LO box is seen as
    Select(dims_in - 1) then Create(dims_out - 1) (diagram.py:453-485):
    ``Create(1) >> ZBox(1, 2) >> W(2)^dagger`` -- the W^dagger sees the replacement
    Create(1, 1) (2 photons) and the Select hides the real Create: 2 + 1 = 3."""
    assert _dims(zw.W(2)) == [("W", [2], [2, 2])]
    assert oc.get_max_dim_for_box(0, oc.core_rec(zw.W(2)), 0, [2], []) == 5
    assert _dims(zw.Create(3) >> zw.W(2)) == [("Create", [], [4]), ("W", [4], [4, 4])]
    got = _dims(zw.Create(1) >> zw.ZBox(1, 2, lambda i: 1) >> zw.W(2).dagger())
    assert got == [("Create", [], [2]), ("ZWZ", [2], [2, 2]), ("W", [2, 2], [3])]
    # a dagger W is capped by the light cone, not by sum(d - 1) + 1: two 1-photon
    # sources, three legs of dim 2 -> min(3 * 1 + 1, 2 + 1) = 3
    got = _dims(zw.Create(1, 1) @ zw.Create(0) >> zw.W(3).dagger())
    assert got[-1] == ("W", [2, 2, 2], [3])
    # bit wires never carry the cone (utils.py:371, 457): a PhotonThresholdDetector's
    # bit output starts no cone, so a later dagger detector is capped at max(0 + 1, 2)
    got = _dims(zw.Create(2) >> diagram.PhotonThresholdDetector()
                >> diagram.PhotonThresholdDetector(is_dagger=True))
    assert got == [("Create", [], [3]), ("PTD", [3], [2]), ("PTD", [2], [2])]


def test_doubled_dims_couple_ket_and_bra_through_classical_wires():
    """In the CP-doubled diagram a measured mode is a Spider(2, 1) on (ket, bra)
    (channel.py:606-615, 638): NON_LO, recorded as Select + Create (diagram.py:453-485)."""
    d = photonic.Create(1) >> photonic.NumberResolvingMeasurement(1)
    dims = _dims(d)
    assert dims[0][0] == "Create" and dims[1][0] == "Create"
    assert dims[-1] == ("Spider", [2, 2], [2])
    assert oracle_eval(d).shape == (2,)


def test_qubit_shortcut_is_the_literal_walk():
    """to_tensor skips the walk for boxes without mode inputs (it returns 2): same network."""
    d = ghz(4, noise=0.9, open_qubits=2)
    a = oracle_eval(d)
    try:
        oc.LITERAL_WALK_ALWAYS = True
        b = oracle_eval(d)
    finally:
        oc.LITERAL_WALK_ALWAYS = False
    assert np.array_equal(a, b)


def test_conjugation_rules_are_the_reference_quirks():
    """SURVEY 8(c) quirks (i)-(iv) as the oracle restates them."""
    r = oc.conj_rec(oc.core_rec(zx.ZBox(1, 1, 0.5 + 0.25j)))
    assert r.p["phase"] == -(0.5 + 0.25j)                       # (i) negated, not conjugated
    r = oc.conj_rec(oc.core_rec(zx.Z(1, 1, 0.25)))
    assert r.p["phase"] == -0.25
    arr = np.array([[1, 2j], [3, 4]])
    box = diagram.Box("A", diagram.Bit(1), diagram.Bit(1), array=arr)
    r = oc.conj_rec(oc.core_rec(box))
    assert np.array_equal(r.p["array"], arr.conj())             # (iii) not transposed
    with pytest.raises(AttributeError):                          # (iv) plain list
        oc.conj_rec(oc.core_rec(zw.ZBox(1, 1, [1, 1j])))
    r = oc.conj_rec(oc.core_rec(zw.ZBox(1, 1, np.array([1, 1j]))))
    assert np.array_equal(r.p["amps"], np.array([1, -1j]))
    r = oc.conj_rec(oc.core_rec(zw.Endo(0.3 + 0.4j)))
    assert r.p["scalar"] == 0.3 - 0.4j


def test_is_pure_and_kraus_follow_channel_py():
    assert oc.is_pure(photonic.Create(1, 1) >> photonic.BS)
    assert not oc.is_pure(lossy_hom(0.8))
    assert not oc.is_pure(qubits.Ket(0) >> qubits.Measure(1))
    assert oc.is_pure(classical.AddN(2))
    for name in sorted(ALL):
        d = ALL[name]()
        assert oc.is_pure(d) == d.is_pure, name
