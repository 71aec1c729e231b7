"""-m gpu: compressed (approximate) contraction on the device (SURVEY.md 8(f) row f4,
reference backends.py:516-545) against the exact CPU oracle and the CPU restatement of
the algorithm (oracle/compressed.py) -- the parity statement the reference's own tests
make for this path is np.allclose with the exact result (test_backends.py:80-134)."""
import numpy as np
import pytest

from cases import EVAL_CASES, compile_channel, oracle_eval, teleportation
from oracle.compressed import contract_compressed as oracle_compressed
from oracle.contract import contract_network
from optyx_b200 import photonic, qubits
from optyx_b200.compressed import contract_compressed
from optyx_b200.core.backends import B200Backend

pytestmark = pytest.mark.gpu

SMALL = ["lossy_hom", "bs_pure", "mzi_chip_pure", "cfg1_4mode", "ghz6_measured", "ghz5_noisy_open3",
         "bitflip", "hom_distinguishable", "cfg3_small", "clifford_t_8x6", "fusion_ii", "fusion_i"]


class HyperCompressedOptimizer:          # recognised by NAME, like cotengra's class
    pass


@pytest.mark.parametrize("name", SMALL)
def test_untruncated_device_compression_matches_exact(name):
    net = compile_channel(EVAL_CASES[name]())
    want = contract_network(net)
    got, info = contract_compressed(net, max_bond=None, cutoff=1e-10, return_info=True)
    got = got.cpu().numpy()
    assert got.shape == want.shape
    assert np.allclose(got, want, atol=1e-8), np.abs(got - want).max()
    assert info["launches"] > 0


def test_reference_compressed_test_circuits_through_the_backend():
    """PURE_CIRCUITS_TO_TEST / MIXED_CIRCUITS_TO_TEST of test_backends.py:39-56, evaluated
    as test_backends.py:82-131 does: backend with a compressed optimiser == exact backend."""
    approx = B200Backend(hyperoptimiser=HyperCompressedOptimizer())
    exact = B200Backend()
    pure = [photonic.BS, photonic.Phase(0.2) @ photonic.Phase(0.3) >> photonic.TBS(0.3),
            photonic.MZI(0.2, 0.8), photonic.seeded_ansatz(4, 4, seed=7)]
    circuits = [photonic.Create(*(1,) * len(c.dom)) >> c for c in pure]
    circuits += [photonic.Create(1, 1) >> photonic.BS >> photonic.Discard(1) @ photonic.qmode
                 >> photonic.NumberResolvingMeasurement(1),
                 qubits.X(0, 1, 0.3) >> teleportation() >> qubits.Measure(1)]
    for d in circuits:
        a = d.eval(approx)
        e = d.eval(exact)
        assert a.tensor.array.shape == e.tensor.array.shape
        assert np.allclose(a.tensor.array, e.tensor.array, atol=1e-8)
        assert np.allclose(a.tensor.array, oracle_eval(d), atol=1e-8)
        pa, pe = a.prob_dist(), e.prob_dist()
        for k in set(pa) | set(pe):
            assert abs(pa.get(k, 0.0) - pe.get(k, 0.0)) <= 1e-5


def test_truncation_follows_the_cpu_restatement():
    """A real truncation (max_bond below the bonds that occur): the error shrinks with
    the cap, vanishes once no bond exceeds it, and at every cap the device result is as
    close to the exact tensor as the CPU restatement of the same algorithm is."""
    net = compile_channel(EVAL_CASES["mzi_chip_pure"]())
    want = contract_network(net)
    _, info = contract_compressed(net, max_bond=None, cutoff=0.0, return_info=True)
    assert info["max_bond_seen"] >= 4
    errs = []
    for chi in (1, 2, 4, 8, info["max_bond_seen"]):
        got = contract_compressed(net, max_bond=chi, cutoff=0.0).cpu().numpy()
        cpu = oracle_compressed(net, max_bond=chi, cutoff=0.0)
        e_dev, e_cpu = float(np.abs(got - want).max()), float(np.abs(cpu - want).max())
        assert e_dev <= 3.0 * e_cpu + 1e-7, (chi, e_dev, e_cpu)
        errs.append(e_dev)
    assert errs[0] > 1e-3 and errs[-1] <= 1e-7 and errs[-2] <= errs[0]


def test_contraction_params_are_forwarded_and_complex64_casts():
    d = EVAL_CASES["cfg1_4mode"]()
    want = oracle_eval(d)
    loose = B200Backend(hyperoptimiser=HyperCompressedOptimizer(), contraction_params={"max_bond": 2})
    tight = B200Backend(hyperoptimiser=HyperCompressedOptimizer(),
                        contraction_params={"max_bond": 64, "cutoff": 1e-12})
    e_loose = np.abs(d.eval(loose).tensor.array - want).max()
    e_tight = np.abs(d.eval(tight).tensor.array - want).max()
    assert e_tight <= 1e-8 and e_loose > e_tight
    c64 = d.eval(B200Backend(hyperoptimiser=HyperCompressedOptimizer(), dtype="complex64"))
    assert c64.tensor.array.dtype == np.complex64
    assert np.abs(c64.tensor.array - want).max() <= 1e-5 * max(1.0, np.abs(want).max())
