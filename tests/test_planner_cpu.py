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


def _rows_histogram(d):
    """Detected-count distribution from ``PermanentBackend._photon_rows`` with the
    ORACLE's permanent, marginalised in Python (what b2_permanent_histogram does)."""
    import itertools
    from math import factorial
    from oracle.permanent import npperm
    from optyx_b200.core.backends import PermanentBackend
    V, where, k = PermanentBackend._photon_rows(d)
    n, M = V.shape
    hist = {}
    for cols in itertools.combinations_with_replacement(range(M), n):
        occ = np.bincount(cols, minlength=M)
        w = abs(npperm(V[:, list(cols)])) ** 2 / np.prod([factorial(int(x)) for x in occ])
        key = [0] * k
        for c in cols:
            if where[c] >= 0:
                key[where[c]] += 1
        hist[tuple(key)] = hist.get(tuple(key), 0.0) + w
    return hist, (V, where, k)


def test_lossy_distinguishable_photon_rows_against_the_tensor_network_oracle():
    """Host half of PermanentBackend.prob_dist_measured: the photon row vectors over
    (spatial mode x internal state + loss environments) and the measured-mode map,
    contracted here with the ORACLE's permanent and marginalised in Python, give
    the same detected-count distribution as the CP-doubled tensor-network oracle."""
    import cases
    from oracle import evalresult as oe
    d = cases.cfg3_interferometer(width=4, n_photons=2, measured=2, inflate=False)
    hist, (V, where, k) = _rows_histogram(d)
    n, M = V.shape
    assert (n, M, k) == (2, 16, 2) and np.allclose((np.abs(V) ** 2).sum(axis=1), 1.0)
    inflated = cases.cfg3_interferometer(width=4, n_photons=2, measured=2)
    want = oe.prob_dist_mixed(cases.oracle_eval(inflated), list(inflated.cod.inside))
    total = sum(want.values())
    assert sum(hist.values()) == pytest.approx(1.0, abs=1e-12)
    assert max(abs(hist.get(key, 0.0) - v / total) for key, v in want.items()) <= 1e-12


def _discard_cases():
    from optyx_b200 import photonic
    from optyx_b200.core.channel import Diagram, mode, qmode
    P, q, cl = photonic, qmode, Diagram.id(mode)
    meas = P.NumberResolvingMeasurement
    u3 = P.Create(1, 1, 1) >> P.seeded_ansatz(3, 3, seed=4)
    return {
        # per-mode discards: must equal one Discard(q ** 2)
        "per_mode_discards": u3 >> meas(1) @ P.Discard(1) @ P.Discard(1),
        "block_discard": u3 >> meas(1) @ P.Discard(2),
        "discard_id_discard": u3 >> P.Discard(1) @ meas(1) @ P.Discard(1),
        # a gate AFTER a discard acts on the wires that are left
        "discard_then_gate": (P.Create(1, 1, 1) >> P.PhotonLoss(0.8) @ q ** 2 >> P.BS @ q
                              >> P.Discard(1) @ q ** 2 >> P.BBS(0.3) >> meas(2)),
        "interleaved": (P.Create(1, 0, 1, 1) >> P.seeded_ansatz(4, 3, seed=2)
                        >> meas(1) @ P.Discard(1) @ q ** 2 >> cl @ P.TBS(0.4)
                        >> cl @ P.Discard(1) @ meas(1)),
        "swap_after_discard": (u3 >> P.Discard(1) @ q ** 2 >> P.Swap(1, 1) >> P.Id(1) @ P.PhotonLoss(0.5)
                               >> meas(2)),
    }


@pytest.mark.parametrize("name", ["per_mode_discards", "block_discard", "discard_id_discard",
                                  "discard_then_gate", "interleaved", "swap_after_discard"])
def test_photon_rows_index_boxes_through_the_live_wire_map(name):
    """A Discard removes its wire from the diagram but not from the output modes: every
    later box must act on the wires that are LEFT (gates / measures after a
    discard hit the wrong columns).  Checked against the CP-doubled oracle."""
    import cases
    from oracle import evalresult as oe
    d = _discard_cases()[name]
    hist, (V, where, k) = _rows_histogram(d)
    assert k == len(d.cod)
    want = oe.prob_dist_mixed(cases.oracle_eval(d), list(d.cod.inside))
    total = sum(want.values())
    assert sum(hist.values()) == pytest.approx(1.0, abs=1e-12)
    keys = set(want) | set(hist)
    assert max(abs(hist.get(key, 0.0) - want.get(key, 0.0) / total) for key in keys) <= 1e-12


def test_photon_rows_per_mode_discards_equal_a_block_discard():
    from optyx_b200.core.backends import PermanentBackend
    c = _discard_cases()
    _, w1, k1 = PermanentBackend._photon_rows(c["per_mode_discards"])
    _, w2, k2 = PermanentBackend._photon_rows(c["block_discard"])
    assert k1 == k2 == 1 and list(w1) == list(w2) == [0, -1, -1]
    _, w3, k3 = PermanentBackend._photon_rows(c["discard_id_discard"])
    assert k3 == 1 and list(w3) == [-1, 0, -1]


def test_leaf_rank_is_unbounded_for_host_valued_leaves():
    """DENSE / VEC leaves are flat on the device (any rank); closed-form
    kinds above the descriptor's 8 axes are built on the host as DENSE."""
    import cases
    from oracle.contract import contract_network
    from optyx_b200 import executor
    from optyx_b200.adapters import network_from_arrays
    from optyx_b200.core import network as nw, zw
    rng = np.random.default_rng(3)
    a = rng.standard_normal((2,) * 10) + 1j * rng.standard_normal((2,) * 10)
    b = rng.standard_normal((2,) * 4)
    net = network_from_arrays([a, b], [tuple("abcdefghij"), tuple("ghij")], tuple("abcdef"))
    net.validate()
    plan = executor.get_plan(net)
    table, _ = executor.leaf_table(net, plan)
    assert table[0]["ndim"] == 1 and table[0]["dims"][0] == 1024
    want = np.einsum("abcdefghij,ghij->abcdef", a, b)
    assert np.allclose(emulate_plan(net, plan)[0], want)
    # a 9-legged W: rank 10 closed form -> DENSE, same values as the oracle's generator
    d = zw.Create(2) >> zw.W(9)
    net = d.to_network()
    w = [l for l in net.leaves if l.name == "W"][0]
    assert w.kind == nw.K_DENSE and len(w.shape) == 10
    assert np.allclose(contract_network(net), cases.oracle_eval(d))
    bad = nw.Network([nw.Leaf(nw.K_SUM, (2,) * 10, tuple(range(10)))], {i: 2 for i in range(10)}, ())
    with pytest.raises(ValueError):
        bad.validate()


# ---- optimiser / contraction-tree intake (backends.py:436-455, 513-545) ---------------------
class _FakeTree:
    def __init__(self, ssa, sliced=()):
        self._ssa, self.sliced_inds = ssa, tuple(sliced)

    def get_ssa_path(self):
        return self._ssa


class HyperOptimizer:                         # recognised by NAME, like cotengra's class
    """Stand-in for cotengra.HyperOptimizer: ``search`` returns a tree with a hand-written
    SSA path (contract the leaves left to right) and one sliced index."""

    def __init__(self, slice_first_closed=False):
        self.slice_first_closed, self.calls = slice_first_closed, []

    def search(self, inputs, output, size_dict):
        self.calls.append((inputs, output, size_dict))
        n = len(inputs)
        ssa = [(0, 1)] + [(n + k - 1, k + 1) for k in range(1, n - 1)]
        sliced = ()
        if self.slice_first_closed:
            closed = [q for ix in inputs for q in ix if q not in output]
            sliced = (closed[0],) if closed else ()
        return _FakeTree(ssa, sliced)


def test_plan_from_hand_written_ssa_path():
    """A given contraction tree (cotengra's tree.get_ssa_path() + sliced_inds) becomes an
    executable plan: checked on the numpy descriptor emulator against the oracle."""
    from optyx_b200 import planner
    rng = np.random.default_rng(5)
    for _ in range(10):
        net = random_network(rng, n_leaves=6, n_inds=8)
        n = len(net.leaves)
        ssa = [(0, 1)] + [(n + k - 1, k + 1) for k in range(1, n - 1)]
        closed = [q for q in net.dims if q not in net.output_inds]
        plan = planner.plan_from_ssa_path(net, ssa, closed[:1])
        assert [tuple(sorted(p)) for p in plan.info.path] == [tuple(sorted(p)) for p in ssa]
        assert plan.nslices == (net.dims[closed[0]] if closed else 1)
        assert np.allclose(emulate_plan(net, plan)[0], contract_network(net))
    with pytest.raises(ValueError):
        planner.plan_from_ssa_path(net, [(0, 0)])
    with pytest.raises(ValueError):
        planner.plan_from_ssa_path(net, ssa, [net.output_inds[0]])


def test_backend_takes_the_tree_the_optimiser_finds():
    from optyx_b200 import executor
    from optyx_b200.core.backends import B200Backend
    from optyx_b200.planner import GivenTree
    rng = np.random.default_rng(6)
    net = random_network(rng, n_leaves=6, n_inds=8)
    opt = HyperOptimizer(slice_first_closed=True)
    backend = B200Backend(hyperoptimiser=opt, contraction_params={"x": 1})
    assert backend.hyperoptimiser is opt and backend.contraction_params == {"x": 1}
    assert backend._optimiser_kind() == "exact"
    ssa, sliced = backend._tree_for(net, "exact")
    inputs, output, size_dict = opt.calls[0]
    assert len(inputs) == len(net.leaves) and all(isinstance(q, str) for ix in inputs for q in ix)
    assert set(size_dict.values()) <= {2, 3, 4}
    assert all(q in net.dims and q not in net.output_inds for q in sliced)
    plan = executor.get_plan(net, ssa_path=ssa, sliced_inds=sliced)
    assert np.allclose(emulate_plan(net, plan)[0], contract_network(net))
    # an explicit tree is accepted as well; anything else is the reference's ValueError
    assert B200Backend(hyperoptimiser=GivenTree(ssa, sliced))._optimiser_kind() == "tree"
    with pytest.raises(ValueError, match="Unsupported hyperoptimiser type"):
        B200Backend(hyperoptimiser="greedy")._optimiser_kind()

    class ReusableHyperCompressedOptimizer:
        pass
    assert B200Backend(hyperoptimiser=ReusableHyperCompressedOptimizer())._optimiser_kind() == "approx"


# ---- the small-step program (csrc/contract_small.cu), emulated ----------------------------
@pytest.mark.parametrize("name", ["lossy_hom", "cfg1_4mode", "ghz8_noisy_measured", "ghz5_noisy_open3",
                                  "cfg3_small", "clifford_t_8x6", "fusion_ii", "hom_distinguishable"])
def test_small_step_program_on_the_emulator(name):
    """Tasks, imports / exports, dependency order, scratch layout and offset tables of
    the persistent small-step launch: the numpy model of the kernel (NaN-poisoned
    scratch and workspace) reproduces the oracle's tensor."""
    import cases
    from optyx_b200 import planner
    d = cases.EVAL_CASES[name]()
    net = cases.compile_channel(d)
    plan = planner.plan_network(net)
    sp = plan.small
    assert sp is not None and len(sp.step_ids) >= planner.SMALL_MIN_STEPS
    assert sum(st.grouped for st in plan.steps) == len(sp.step_ids)
    assert sp.smem_bytes(16) <= 226 * 1024 and sp.smem_bytes(8) <= sp.smem_bytes(16)
    got = emulate_plan(net, plan)[0]
    want = contract_network(net)
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12)


def test_small_step_program_batched_and_sliced():
    import cases
    from optyx_b200 import planner
    from optyx_b200.core import network as nw
    # a batch: one chain of tasks per eval, per-eval strides on every import / export
    nets = [cases.compile_channel(cases.fusion_teleport_input(0.1 * (j + 1))) for j in range(3)]
    for net in nets:
        net.leaves.append(nw.Leaf(nw.K_DENSE, (), (), 0, np.array([net.scalar], dtype=complex), "scalar"))
        net.scalar = 1.0 + 0j
    assert nets[0].structure_key() == nets[1].structure_key()
    plan = planner.plan_network(nets[0], batch=3)
    assert plan.small is not None
    got = emulate_plan(nets[0], plan, nets=nets)
    for e, net in enumerate(nets):
        assert np.allclose(got[e], contract_network(net), rtol=1e-12, atol=1e-12)
    # a sliced plan: only slice-invariant steps enter the program
    net = cases.compile_channel(cases.random_clifford_t(12, 10, seed=3))
    plan = planner.plan_network(net, target_size=16.0)
    assert plan.nslices > 1 and plan.small is not None
    assert all(not plan.steps[k].per_slice for k in plan.small.step_ids)
    assert np.allclose(emulate_plan(net, plan)[0], contract_network(net), rtol=1e-12, atol=1e-12)


def test_small_step_tasks_run_in_parallel_for_one_eval_and_in_one_chain_per_batch_eval():
    import cases
    from optyx_b200 import planner
    net = cases.compile_channel(cases.ghz(12, open_qubits=3, noise=0.9))
    single = planner.plan_network(net).small
    batched = planner.plan_network(net, batch=4).small
    assert single.n_tasks > batched.n_tasks >= 1
    steps_of = [int(single.blob[single.blob_off[t] + 3]) for t in range(single.n_tasks)]
    assert max(steps_of) <= planner.SMALL_TASK_STEPS


# ---- native (C++) path search behind the C ABI ---------------------------------------------
def test_native_greedy_and_cost_agree_with_the_python_planner():
    """b2_plan_greedy gives a valid SSA pair order; b2_plan_cost == analyse_path on it
    (MACs, widest intermediate, slice count and roofline estimate), sliced or not."""
    import cases
    from optyx_b200 import planner
    net = cases.compile_channel(cases.random_clifford_t(12, 10, seed=4))
    inputs = [frozenset(l.inds) for l in net.leaves]
    output = frozenset(net.output_inds)
    nn = planner.NativeNetwork(inputs, net.dims, output)
    n = len(inputs)
    for seed, temp in ((0, 0.0), (1, 0.5), (2, 1.0)):
        p = nn.greedy(seed=seed, temperature=temp, costmod=1.0)
        assert p.shape == (n - 1, 2)
        used = set()
        for k, (a, b) in enumerate(p.tolist()):
            assert a != b and a < n + k and b < n + k and a not in used and b not in used
            used.update((a, b))
        path = [tuple(x) for x in p.tolist()]
        closed = sorted({q for s in inputs for q in s} - output)
        for sliced in ((), tuple(closed[:3])):
            info = planner.analyse_path(inputs, net.dims, output, path, sliced)
            macs, peak, nslices, est = nn.cost(p, sliced)
            assert (macs, peak, nslices) == (info.macs, info.peak, info.nslices)
            assert abs(est - info.time_est) <= 1e-12 * max(1.0, est)
    # same seed, same tree (every rank of a job must find the same plan)
    assert np.array_equal(nn.greedy(seed=5, temperature=0.7), nn.greedy(seed=5, temperature=0.7))
    # a plan built from the native search contracts to the oracle's tensor
    plan = planner.plan_network(net, target_size=64.0, min_slices=8)
    assert plan.nslices >= 8 and plan.info.peak <= 64
    assert np.allclose(emulate_plan(net, plan)[0], contract_network(net), rtol=1e-12, atol=1e-12)


def test_min_slices_keeps_slicing_with_the_cheapest_indices():
    from optyx_b200 import planner
    rng = np.random.default_rng(11)
    net = random_network(rng, n_leaves=9, n_inds=12, n_out=1, dims=(2,))
    inputs = [frozenset(l.inds) for l in net.leaves]
    output = frozenset(net.output_inds)
    base = planner.find_path(inputs, net.dims, output, trials=4)
    more = planner.find_path(inputs, net.dims, output, trials=4, min_slices=4)
    assert base.nslices == 1 and more.nslices >= 4
    assert not (set(more.sliced) & output)
    plan = planner.plan_network(net, min_slices=4)
    assert np.allclose(emulate_plan(net, plan)[0], contract_network(net), rtol=1e-12, atol=1e-12)


def test_array_valued_parameters_compile_a_whole_batch_in_one_pass():
    """eval_params: a diagram built with ARRAY phases / amplitudes compiles once into
    the [batch, P] matrix of leaf values -- identical to compiling every eval on its own."""
    import math
    import cases
    from optyx_b200.executor import params_vector
    th = np.linspace(0.0, math.pi / 2, 6)
    both = [
        (lambda: cases.fusion_link((np.cos(th), np.sin(th)), "bell"),
         lambda j: cases.fusion_link((math.cos(th[j]), math.sin(th[j])), "bell")),
        (lambda: cases.fusion_teleport_input(np.arange(6) / 6), lambda j: cases.fusion_teleport_input(j / 6)),
    ]
    for batched, single in both:
        net = cases.compile_channel(batched())
        assert net.batch == 6
        pm = net.params_matrix(6)
        for j in range(6):
            one = cases.compile_channel(single(j))
            assert one.structure_key() == net.structure_key() and one.batch is None
            assert np.abs(pm[j] - params_vector(one)).max() <= 1e-15
            sc = net.scalar[j] if np.ndim(net.scalar) else net.scalar
            assert abs(sc - one.scalar) <= 1e-15


def test_subtree_reconfiguration_never_hurts_and_keeps_the_tensor():
    """b2_plan_reconfigure: optimal pair order inside every subtree of <= 8 frontier
    nodes.  Cost never goes up, the rewritten SSA path is valid, and the plan compiled
    from it contracts to the oracle's tensor (hyper-indices, open outputs, slicing)."""
    import cases
    from optyx_b200 import planner
    for d, sliced_n in ((cases.random_clifford_t(12, 10, seed=4), 2), (cases.ghz(8, open_qubits=3, noise=0.9), 0),
                        (cases.cfg1_interferometer(4, (1, 1, 0, 0)), 0)):
        net = cases.compile_channel(d)
        inputs = [frozenset(l.inds) for l in net.leaves]
        output = frozenset(net.output_inds)
        nn = planner.NativeNetwork(inputs, net.dims, output)
        closed = sorted({q for s in inputs for q in s} - output)
        sliced = tuple(closed[:sliced_n])
        for seed in (1, 2, 3):
            p = nn.greedy(seed=seed, temperature=1.0)
            before = nn.cost(p, sliced)
            p2 = nn.reconfigure(p, sliced, max_leaves=8)
            after = nn.cost(p2, sliced)
            assert after[0] <= before[0] * (1 + 1e-12)
            n = len(inputs)
            used = set()
            for k, (a, b) in enumerate(p2.tolist()):
                assert a != b and a < n + k and b < n + k and a not in used and b not in used
                used.update((a, b))
            plan = planner.plan_from_ssa_path(net, [tuple(x) for x in p2.tolist()], sliced)
            assert np.allclose(emulate_plan(net, plan)[0], contract_network(net), rtol=1e-12, atol=1e-12)


def test_threaded_native_search_does_not_depend_on_the_thread_count(monkeypatch):
    """The C++ trials, the subtree reconfigurations and the slicing of the candidate trees run
    on a thread pool; every rank of a sliced job must still find the SAME plan whatever its
    core count: same pair order, same sliced indices with 1 thread and with many."""
    import os
    import cases
    from optyx_b200 import planner
    net = cases.compile_channel(cases.random_clifford_t(12, 10, seed=4))
    inputs = [frozenset(l.inds) for l in net.leaves]
    output = frozenset(net.output_inds)
    assert len(inputs) >= planner.NATIVE_MIN_TENSORS
    many = planner.find_path(inputs, net.dims, output, trials=2, seed=3, target_size=64.0, min_slices=8)
    monkeypatch.setattr(os, "cpu_count", lambda: 1)
    one = planner.find_path(inputs, net.dims, output, trials=2, seed=3, target_size=64.0, min_slices=8)
    assert many.path == one.path and tuple(many.sliced) == tuple(one.sliced)
    assert many.nslices >= 8 and many.peak <= 64


def test_descriptor_table_is_cached_per_plan_and_values_follow_the_network():
    """executor.leaf_table: the descriptor table depends on the network STRUCTURE only and is
    built once per plan; the host values are gathered from whichever network is passed --
    a structurally identical network with other parameters must get ITS values."""
    import cases
    from optyx_b200 import executor
    from optyx_b200.core import network as nw
    net_a = cases.compile_channel(cases.lossy_hom(0.8))
    net_b = cases.compile_channel(cases.lossy_hom(0.3))
    assert net_a.structure_key() == net_b.structure_key()
    plan = plan_network(net_a)
    t1, p_a = executor.leaf_table(net_a, plan)
    t2, p_b = executor.leaf_table(net_b, plan)
    assert t2 is t1 and not t1.flags.writeable            # one table per plan
    want_a = executor.params_vector(net_a)
    want_b = executor.params_vector(net_b)
    assert np.array_equal(p_a, want_a) and np.array_equal(p_b, want_b)
    assert not np.array_equal(p_a, p_b)                    # the loss parameter differs
    # the cached table equals a from-scratch build
    fresh = plan_network(net_b)
    t3, _ = executor.leaf_table(net_b, fresh)
    assert t3 is not t1 and t3.tobytes() == t1.tobytes()
    host = [k for k, l in enumerate(net_a.leaves) if l.kind in (nw.K_VEC, nw.K_DENSE)]
    assert len(host) > 0
