"""CPU: the oracle's EvalResult restatement against the reference's own
literal-array tests (test/test_backends.py:255-405, 456-465), plus the product's
EvalResult host logic (dict conversion, normalisation, exceptions)."""
import numpy as np
import pytest

from oracle import evalresult as ev


def test_prob_lookup_and_missing_key():
    v = np.array([1 / np.sqrt(2), 1j / np.sqrt(2)], dtype=complex)
    p = ev.prob_dist(v, ["mode"], ev.AMP)
    assert p[(0,)] == pytest.approx(0.5, rel=1e-12) and p[(1,)] == pytest.approx(0.5, rel=1e-12)
    assert p.get((2,), 0.0) == 0.0


def test_mixed_interleaved_partial_trace_numeric_hygiene():
    dm = np.zeros((2, 2, 2), dtype=complex)
    dm[0, 0, 0] = 0.499999999999 + 1e-14j
    dm[0, 1, 1] = 0.100000000001 - 1e-14j
    dm[1, 0, 0] = 0.2000000000005 + 1e-14j
    dm[1, 1, 1] = 0.2000000000005 - 1e-14j
    dm[0, 0, 1] = -1e-13
    dm[1, 1, 0] = -1e-13
    probs = ev.prob_dist(dm, ["mode", "other"], ev.DM)
    assert probs[(0,)] == pytest.approx(0.6, rel=1e-12)
    assert probs[(1,)] == pytest.approx(0.4, rel=1e-12)


def test_prob_branch_rounding_and_sum():
    P = np.array([[1 / 3, 1 / 6], [1 / 6, 1 / 3]], dtype=float)
    d = ev.prob_dist(P, ["mode", "mode"], ev.PROB, round_digits=4)
    assert sum(d.values()) == pytest.approx(1.0, rel=1e-12)
    assert d[(0, 0)] == round(float(P[0, 0]), 4) and d[(1, 1)] == round(float(P[1, 1]), 4)


def test_mixed_simple_single_measured_single_unmeasured():
    dm = np.zeros((2, 2, 2), dtype=complex)
    dm[0, 0, 0] = 0.6 + 1e-14j
    dm[1, 1, 1] = 0.4 - 1e-14j
    dm[0, 1, 0] = -1e-13
    dm[1, 0, 1] = -1e-13
    probs = ev.prob_dist(dm, ["mode", "qubit"], ev.DM)
    assert probs[(0,)] == pytest.approx(0.6, rel=1e-12)
    assert probs[(1,)] == pytest.approx(0.4, rel=1e-12)
    assert sum(probs.values()) == pytest.approx(1.0, rel=1e-12)


def test_mixed_two_measured_one_unmeasured_matches_manual_trace():
    P = np.array([[0.1, 0.2], [0.3, 0.4]])
    W = np.array([0.6, 0.4])
    dm = np.zeros((2, 2, 2, 2), dtype=complex)
    for k in range(2):
        for l in range(2):
            dm[k, l, 0, 0] = P[k, l] * W[0]
            dm[k, l, 1, 1] = P[k, l] * W[1]
    dm[0, 0, 0, 1] = 1e-13
    dm[1, 1, 1, 0] = -1e-13
    probs = ev.prob_dist(dm, ["mode", "bit", "qubit"], ev.DM)
    manual = {(k, l): (dm[k, l, 0, 0].real + dm[k, l, 1, 1].real) for k in range(2)
              for l in range(2)}
    Z = sum(manual.values())
    for k in manual:
        assert probs[k] == pytest.approx(manual[k] / Z, rel=1e-12)


def test_mixed_round_digits_applied_after_normalisation():
    dm = np.zeros((2, 2, 2), dtype=complex)
    dm[0, 0, 0] = 0.333333333333
    dm[1, 1, 1] = 0.666666666667
    d = ev.prob_dist(dm, ["mode", "qubit"], ev.DM, round_digits=6)
    a, b = round(0.333333333333, 6), round(0.666666666667, 6)
    assert d[(0,)] == pytest.approx(a / (a + b), rel=1e-12)
    assert d[(1,)] == pytest.approx(b / (a + b), rel=1e-12)


def test_mixed_requires_at_least_one_measured_type():
    dm = np.eye(4, dtype=complex).reshape(2, 2, 2, 2) / 4.0
    with pytest.raises(ValueError):
        ev.prob_dist(dm, ["qubit", "qmode"], ev.DM)


def test_mixed_two_unmeasured_pairs_partial_trace():
    dm = np.zeros((2, 2, 2, 2, 2), dtype=float)
    w_m, w_u1, w_u2 = {0: 0.55, 1: 0.45}, np.array([0.6, 0.4]), np.array([0.7, 0.3])
    for m in (0, 1):
        for r1 in (0, 1):
            for r2 in (0, 1):
                dm[m, r1, r1, r2, r2] = w_m[m] * w_u1[r1] * w_u2[r2]
    dm[0, 0, 1, 0, 0] = 1e-13
    dm[1, 1, 0, 1, 0] = -1e-13
    probs = ev.prob_dist(dm, ["mode", "x", "y"], ev.DM)
    assert probs[(0,)] == pytest.approx(0.55, rel=1e-12)
    assert probs[(1,)] == pytest.approx(0.45, rel=1e-12)


def test_mixed_all_off_diagonal_mass_raises_zero_total():
    dm = np.zeros((2, 2, 2))
    dm[0, 0, 1] = 0.6
    dm[1, 1, 0] = 0.4
    with pytest.raises(ValueError):
        ev.prob_dist(dm, ["mode", "x"], ev.DM)


def test_not_a_state_raises():
    with pytest.raises(ValueError):
        ev.prob_dist(np.ones((2, 2)), ["mode"], ev.AMP, n_dom=1)
    with pytest.raises(TypeError):
        ev.amplitudes(np.ones(2), ev.DM)


def test_product_evalresult_host_logic():
    """The product's EvalResult: density matrix from AMP (backends.py:147-167,
    test_backends.py:267-274, 309-320), exceptions, dict conversion."""
    from optyx_b200.core.backends import EvalResult, ResultTensor, StateType
    v = np.array([np.exp(1j * 0.2) / np.sqrt(3), np.exp(-1j * 0.3) * np.sqrt(2 / 3)])
    r = EvalResult(ResultTensor((), (2,), v), ("mode",), StateType.AMP)
    dm = r.density_matrix
    assert dm.shape == (2, 2) and np.allclose(dm, np.outer(v.conj(), v), atol=1e-12)
    amps = r.amplitudes()
    assert np.isclose(sum(abs(a) ** 2 for a in amps.values()), 1.0)
    dmr = EvalResult(ResultTensor((), (2, 2), np.array([[0.7, 0.1j], [-0.1j, 0.3]])),
                     ("mode",), StateType.DM)
    assert np.allclose(dmr.density_matrix, [[0.7, 0.1j], [-0.1j, 0.3]])
    with pytest.raises(TypeError):
        dmr.amplitudes()
    bad = EvalResult(ResultTensor((2,), (2,), np.eye(2)), ("mode",), StateType.AMP)
    with pytest.raises(ValueError):
        bad.prob_dist()
    with pytest.raises(ValueError):
        _ = bad.density_matrix


def test_index_tuples_match_unravel_index():
    """Host side of prob_dist: keys of the selected flat positions, C order, as tuples
    of Python ints (dense selections walk itertools.product, sparse ones unravel)."""
    from optyx_b200.core.backends import _index_tuples
    rng = np.random.default_rng(0)
    for shape in [(2, 3, 4), (5,), (2,) * 10, (7, 7, 3), (1,), (4, 1, 2)]:
        n = int(np.prod(shape))
        for dens in (1.0, 0.5, 0.05, 0.0):
            mask = (rng.random(n) < dens).astype(np.uint8)
            flat = np.flatnonzero(mask)
            got = list(_index_tuples(flat, mask, shape))
            want = [tuple(int(x) for x in np.unravel_index(f, shape)) for f in flat]
            assert got == want, (shape, dens)
            assert all(type(i) is int for key in got for i in key)
