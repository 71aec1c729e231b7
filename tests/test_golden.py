"""Committed fixtures (tests/golden/, generator: tests/golden/make_golden.py).

* ``reference_known_answers.json``: literal numbers from the reference's own tests /
  docstrings -- the oracle (host compile + numpy contraction) must reproduce them (CPU),
  and so must the CUDA path (``-m gpu``).
* ``eval_cases.npz``: oracle result tensors of the named eval cases -- a regression pin for
  the oracle + host compiler on CPU, and committed vectors for the CUDA path on the GPU box."""
import json
import os

import numpy as np
import pytest

from cases import EVAL_CASES, lossy_hom, oracle_eval
from oracle import evalresult as ev
from optyx_b200 import photonic, qubits
from optyx_b200.core import diagram, zw
from optyx_b200.core.diagram import Id, Mode, Swap

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KNOWN = json.load(open(os.path.join(GOLDEN, "reference_known_answers.json")))
VECTORS = np.load(os.path.join(GOLDEN, "eval_cases.npz"))


def _hom(s1, s2):
    zb_i = zw.ZBox(1, 1, lambda i: (np.sin(0.25 * np.pi) * 1j) ** i)
    zb_1 = zw.ZBox(1, 1, lambda i: (np.cos(0.25 * np.pi)) ** i)
    bs = (zw.W(2) @ zw.W(2) >> zb_i @ zb_1 @ zb_1 @ zb_i
          >> Id(Mode(1)) @ Swap(diagram.mode, diagram.mode) @ Id(Mode(1))
          >> zw.W(2).dagger() @ zw.W(2).dagger())
    return zw.Create(1) @ zw.Create(1) >> bs >> zw.Select(s1) @ zw.Select(s2)


def test_fixture_files_cover_every_case():
    assert sorted(VECTORS.files) == sorted(EVAL_CASES)
    for entry in KNOWN.values():
        assert entry["cite"] and "values" in entry


def test_oracle_reproduces_reference_known_answers():
    for key, (re, im) in KNOWN["hom_amplitudes"]["values"].items():
        s1, s2 = (int(x) for x in key.split(","))
        assert abs(complex(oracle_eval(_hom(s1, s2))) - complex(re, im)) <= 1e-8
    p = ev.prob_dist(oracle_eval(lossy_hom(0.8)), ["mode"], ev.DM)
    for key, val in KNOWN["lossy_hom_prob"]["values"].items():
        assert p[(int(key),)] == pytest.approx(val, abs=1e-12)
    p = ev.prob_dist(oracle_eval(photonic.Create(1, 1) >> photonic.BS), ["qmode", "qmode"], ev.AMP)
    assert round(p[(2, 0)], 1) == KNOWN["bs_single_prob"]["values"]["2,0"]
    assert np.allclose(oracle_eval(qubits.Z(0, 1) >> qubits.BitFlipError(0.43)),
                       np.array(KNOWN["bit_flip_error"]["values"]))
    assert np.allclose(oracle_eval(qubits.Z(0, 1) >> qubits.DephasingError(0.43)),
                       np.array(KNOWN["dephasing_error"]["values"]))
    w = KNOWN["mixed_prob_dist_literal"]["values"]
    rho = np.zeros((2, 2, 2), dtype=complex)          # (measured, ket, bra), test_backends.py:383-405
    for m, weight in w.items():
        rho[int(m)] = weight * np.eye(2) / 2
    p = ev.prob_dist(rho, ["bit", "qubit"], ev.DM)
    for m, weight in w.items():
        assert p[(int(m),)] == pytest.approx(weight, rel=1e-12)


@pytest.mark.parametrize("name", sorted(EVAL_CASES))
def test_oracle_matches_committed_vectors(name):
    got = np.asarray(oracle_eval(EVAL_CASES[name]()))
    want = VECTORS[name]
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max())


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(EVAL_CASES))
def test_cuda_path_matches_committed_vectors(name):
    from optyx_b200.core.backends import B200Backend
    want = VECTORS[name]
    got = EVAL_CASES[name]().eval(B200Backend()).tensor.array
    assert np.shape(got) == want.shape                                   # exact bookkeeping
    assert np.abs(np.asarray(got) - want).max() <= 1e-10 * max(1.0, np.abs(want).max())
    if name in ("lossy_hom", "cfg1_4mode", "ghz6_measured", "clifford_t_8x6", "cfg3_small"):
        got64 = EVAL_CASES[name]().eval(B200Backend(dtype="complex64")).tensor.array
        assert np.abs(np.asarray(got64) - want).max() <= 1e-5 * max(1.0, np.abs(want).max())


@pytest.mark.gpu
def test_cuda_path_reproduces_reference_known_answers():
    from optyx_b200.core.backends import B200Backend
    p = lossy_hom(0.8).eval(B200Backend()).prob_dist()
    for key, val in KNOWN["lossy_hom_prob"]["values"].items():
        assert p[(int(key),)] == pytest.approx(val, abs=1e-12)
    r = (photonic.Create(1, 1) >> photonic.BS).eval(B200Backend())
    assert round(r.single_prob((2, 0)), 1) == KNOWN["bs_single_prob"]["values"]["2,0"]
    for name, gate in (("bit_flip_error", qubits.BitFlipError), ("dephasing_error", qubits.DephasingError)):
        a = (qubits.Z(0, 1) >> gate(0.43)).eval(B200Backend()).tensor.array
        assert np.allclose(a, np.array(KNOWN[name]["values"]))
