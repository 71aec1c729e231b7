"""CPU: planner / layouts / slicing / descriptors, via the numpy emulator of the
kernel descriptor semantics, against the oracle on random hyper-networks."""
import numpy as np
import pytest

from helpers import emulate_plan, random_network
from oracle.contract import contract_network
from optyx_b200.planner import plan_network


@pytest.mark.parametrize("seed", range(60))
def test_plan_emulation_matches_oracle(seed):
    rng = np.random.default_rng(seed)
    net = random_network(rng, n_leaves=int(rng.integers(1, 10)), n_inds=int(rng.integers(3, 12)),
                         n_out=int(rng.integers(0, 4)), repeat_in_leaf=seed % 3 == 0)
    if seed % 5 == 0:
        net.output_inds = net.output_inds + net.output_inds[:1]   # repeated open index
    want = contract_network(net)
    other = contract_network(net, order="random", seed=seed)
    assert np.allclose(want, other, rtol=1e-10, atol=1e-10)       # two oracle orders agree
    for thr in [(16, 16, 4), (1, 1, 1)]:
        for tgt in [None, 8]:
            plan = plan_network(net, target_size=tgt, gemm_threshold=thr, trials=3)
            got = emulate_plan(net, plan)[0]
            assert got.shape == want.shape
            assert np.allclose(got, want, rtol=1e-9, atol=1e-9), (seed, thr, tgt)


@pytest.mark.parametrize("batch", [2, 4])
def test_batched_plan_emulation(batch):
    nets = []
    for e in range(batch):
        net = random_network(np.random.default_rng(7), n_leaves=8, n_inds=9, n_out=2)
        prng = np.random.default_rng(1000 + e)
        for leaf in net.leaves:
            if leaf.params is not None:
                leaf.params = prng.standard_normal(leaf.size) + 1j * prng.standard_normal(leaf.size)
        nets.append(net)
    for thr in [(16, 16, 4), (1, 1, 1)]:
        plan = plan_network(nets[0], batch=batch, gemm_threshold=thr)
        got = emulate_plan(nets[0], plan, nets=nets)
        for e, net in enumerate(nets):
            assert np.allclose(got[e], contract_network(net), rtol=1e-9, atol=1e-9)


def test_slices_partition_across_ranks():
    net = random_network(np.random.default_rng(3), n_leaves=9, n_inds=10, n_out=1)
    plan = plan_network(net, target_size=4)
    assert plan.nslices > 1
    want = contract_network(net)
    half = plan.nslices // 2
    part = emulate_plan(net, plan, slice_range=(0, half))[0] + \
        emulate_plan(net, plan, slice_range=(half, plan.nslices))[0]
    # invariant steps do not depend on the slice range; only the accumulated output adds
    assert np.allclose(part, want, rtol=1e-9, atol=1e-9)


def test_path_candidates_are_compared_after_slicing():
    """With a target size the search keeps the tree whose SLICED contraction has the
    best roofline estimate, which is never worse than slicing the unsliced winner."""
    import random as _random
    from optyx_b200 import planner
    rnd = _random.Random(5)
    n_t, n_i = 40, 60
    dims = {i: 2 for i in range(n_i)}
    inputs = []
    for _ in range(n_t):
        inputs.append(frozenset(rnd.sample(range(n_i), 4)))
    used = set().union(*inputs)
    # close every index that appears once by giving it a partner tensor
    count = {q: sum(q in s for s in inputs) for q in used}
    inputs += [frozenset([q]) for q, c in count.items() if c == 1]
    output = frozenset()
    base = planner.find_path(inputs, dims, output, trials=12, seed=3)
    target = max(4.0, base.peak / 16)
    old_way = planner.slice_path(inputs, dims, output, base, target)
    new_way = planner.find_path(inputs, dims, output, trials=12, seed=3, target_size=target,
                                slice_candidates=6)
    assert new_way.peak <= target + 1e-9
    assert new_way.time_est <= old_way.time_est * (1 + 1e-12)
    assert new_way.nslices >= 1 and new_way.time_est > 0


def test_quimb_adapter_on_stand_in_objects():
    """INTEGRATION.md adapter path: quimb-like tensors (data + index names, a hyper
    index shared by three tensors, an integer-valued tensor) -> Network -> oracle."""
    from types import SimpleNamespace
    from optyx_b200.adapters import network_from_quimb
    from oracle.contract import contract_network
    rng = np.random.default_rng(3)
    a = rng.standard_normal((2, 3, 4)) + 1j * rng.standard_normal((2, 3, 4))
    b = rng.standard_normal((4, 3, 5)) + 1j * rng.standard_normal((4, 3, 5))
    c = np.arange(6).reshape(3, 2)                   # int dtype, shares the hyper index "h"
    tn = SimpleNamespace(
        tensors=[SimpleNamespace(data=a, inds=("i", "h", "k")),
                 SimpleNamespace(data=b, inds=("k", "h", "o")),
                 SimpleNamespace(data=c, inds=("h", "p"))],
        outer_inds=lambda: ("i", "o", "p"))
    net = network_from_quimb(tn)
    assert net.output_shape == (2, 5, 2) and all(l.params.dtype == np.complex128 for l in net.leaves)
    want = np.einsum("ihk,kho,hp->iop", a, b, c)
    assert np.allclose(contract_network(net), want, rtol=1e-12, atol=1e-12)
    with pytest.raises(ValueError):
        network_from_quimb(tn, output_inds=("i", "zzz"))


def test_permanent_backend_host_logic():
    """PermanentBackend: single-particle matrix of Create >> gates,
    reference error behaviour for non-LO circuits (test_backends.py:470-479)."""
    from optyx_b200 import photonic
    from optyx_b200.core.backends import PermanentBackend
    d = photonic.Create(1, 0) @ photonic.Create(2) >> photonic.seeded_ansatz(3, 3, seed=1) \
        >> photonic.Swap(1, 1) @ photonic.Id(1)
    T, creations = PermanentBackend._single_particle(d)
    assert creations == [1, 0, 2] and T.shape == (3, 3)
    assert np.allclose(T @ T.conj().T, np.eye(3), atol=1e-12)
    with pytest.raises(AssertionError):
        PermanentBackend._single_particle(d >> photonic.NumberResolvingMeasurement(3))
    with pytest.raises(AssertionError):
        PermanentBackend._single_particle(photonic.Create(1) >> photonic.PhotonLoss(0.5))
    with pytest.raises(NotImplementedError):
        PermanentBackend._single_particle(photonic.BS)


def test_lossy_distinguishable_photon_rows_against_the_tensor_network_oracle():
    """Host half of PermanentBackend.prob_dist_measured: the photon row vectors over
    (spatial mode x internal state + loss environments) and the measured-mode map,
    contracted here with the ORACLE's permanent and marginalised in Python, give
    the same detected-count distribution as the CP-doubled tensor-network oracle."""
    import itertools
    from math import factorial
    import cases
    from oracle import evalresult as oe
    from oracle.permanent import npperm
    from optyx_b200.core.backends import PermanentBackend
    d = cases.cfg3_interferometer(width=4, n_photons=2, measured=2, inflate=False)
    V, where, k = PermanentBackend._photon_rows(d)
    n, M = V.shape
    assert (n, M, k) == (2, 16, 2) and np.allclose((np.abs(V) ** 2).sum(axis=1), 1.0)
    hist = {}
    for cols in itertools.combinations_with_replacement(range(M), n):
        occ = np.bincount(cols, minlength=M)
        w = abs(npperm(V[:, list(cols)])) ** 2 / np.prod([factorial(int(x)) for x in occ])
        key = [0] * k
        for c in cols:
            if where[c] >= 0:
                key[where[c]] += 1
        hist[tuple(key)] = hist.get(tuple(key), 0.0) + w
    inflated = cases.cfg3_interferometer(width=4, n_photons=2, measured=2)
    want = oe.prob_dist_mixed(cases.oracle_eval(inflated), list(inflated.cod.inside))
    total = sum(want.values())
    assert sum(hist.values()) == pytest.approx(1.0, abs=1e-12)
    assert max(abs(hist.get(key, 0.0) - v / total) for key, v in want.items()) <= 1e-12
