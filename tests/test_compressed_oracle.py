"""CPU: the compressed-contraction oracle (oracle/compressed.py, the checker for SURVEY.md
8(f) row f4) against the exact oracle -- the parity statement the reference's own tests make
for its approximate path (test/test_backends.py:80-134: compressed == exact within
np.allclose on small circuits), plus the truncation properties of the algorithm."""
import numpy as np
import pytest

from cases import EVAL_CASES, compile_channel, teleportation
from oracle.compressed import contract_compressed
from oracle.contract import contract_network
from optyx_b200 import photonic, qubits

SMALL = ["lossy_hom", "bs_pure", "mzi_chip_pure", "cfg1_4mode", "ghz6_measured", "ghz5_noisy_open3",
         "bitflip", "hom_distinguishable", "cfg3_small", "clifford_t_8x6", "fusion_ii", "fusion_i"]


@pytest.mark.parametrize("name", SMALL)
def test_untruncated_compressed_contraction_is_exact(name):
    net = compile_channel(EVAL_CASES[name]())
    want = contract_network(net)
    got = contract_compressed(net, max_bond=None, cutoff=0.0)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 1e-10 * max(1.0, np.abs(want).max())
    # the reference's default (cutoff 1e-10, no bond cap): np.allclose, as its tests assert
    got = contract_compressed(net, max_bond=None, cutoff=1e-10)
    assert np.allclose(got, want, atol=1e-8)


def test_reference_compressed_test_circuits():
    """PURE_CIRCUITS_TO_TEST / MIXED_CIRCUITS_TO_TEST of test_backends.py:39-56."""
    pure = [photonic.BS, photonic.Phase(0.2) @ photonic.Phase(0.3) >> photonic.TBS(0.3),
            photonic.MZI(0.2, 0.8), photonic.seeded_ansatz(4, 4, seed=7)]
    for circ in pure:
        d = photonic.Create(*(1,) * len(circ.dom)) >> circ
        net = compile_channel(d)
        assert np.allclose(contract_compressed(net), contract_network(net), atol=1e-8)
    mixed = [photonic.Create(1, 1) >> photonic.BS >> photonic.Discard(1) @ photonic.qmode
             >> photonic.NumberResolvingMeasurement(1),
             qubits.X(0, 1, 0.3) >> teleportation() >> qubits.Measure(1)]
    for d in mixed:
        net = compile_channel(d)
        assert np.allclose(contract_compressed(net), contract_network(net), atol=1e-8)


def test_truncation_error_shrinks_with_the_bond_cap():
    net = compile_channel(EVAL_CASES["mzi_chip_pure"]())
    want = contract_network(net)
    _, info = contract_compressed(net, max_bond=None, cutoff=0.0, return_info=True)
    assert info["max_bond_seen"] >= 4
    errs = []
    for chi in (1, 2, 4, 8, info["max_bond_seen"]):
        got, inf = contract_compressed(net, max_bond=chi, cutoff=0.0, return_info=True)
        assert got.shape == want.shape and inf["compressions"] > 0
        errs.append(float(np.abs(got - want).max()))
    assert errs[0] > 1e-3                       # chi = 1 really truncates
    assert errs[-1] <= 1e-10                    # a cap no bond exceeds: exact
    assert errs[-2] <= errs[0]                  # and the error goes down on the way
