"""-m gpu: ``diagram.eval(B200Backend())`` end to end (leaf build -> plan ->
CUDA contractions -> prob kernels, all through the C ABI) against the oracle
and the reference's known answers."""
import numpy as np
import pytest

from cases import (cfg1_interferometer, cfg3_interferometer, compile_channel, distinguishable_bs,
                   fusion_i_spider, fusion_ii_teleportation, ghz, lossy_hom, oracle_eval,
                   random_clifford_t, statevector_clifford_t)
from oracle import evalresult as ev
from oracle.contract import contract_network
from optyx_b200 import classical, photonic, qubits
from optyx_b200.core.backends import B200Backend, EvalResult, ResultTensor, StateType
from optyx_b200.core.channel import bit, qubit

pytestmark = pytest.mark.gpu

TOL128, TOL64 = 1e-10, 1e-5


def _close(got, want, tol):
    scale = max(1.0, float(np.abs(want).max())) if np.size(want) else 1.0
    return np.abs(np.asarray(got) - np.asarray(want)).max() <= tol * scale


def _types(d):
    return list(d.cod.inside)


def _check_prob_dist(res, arr, d):
    """Piece (4) is EXACT given the tensor: the device's prob_dist of its own result has
    the key set and values the oracle's EvalResult restatement gives for that same
    array (which entries are exact zeros -- hence dict keys, backends.py:274-289 --
    depends on the contraction order, for quimb as for us).  Against the oracle's own
    tensor the probabilities agree to 1e-10, keys of negligible weight aside."""
    st = ev.AMP if res.state_type is StateType.AMP else ev.DM
    got = {tuple(int(i) for i in k): v for k, v in res.prob_dist().items()}
    same = ev.prob_dist(res.tensor.array, _types(d), st)
    assert set(got) == set(same)                                       # exact key set
    for k, v in got.items():
        assert abs(v - same[k]) <= 1e-12
    want = ev.prob_dist(arr, _types(d), st)
    for k in set(got) | set(want):
        assert abs(got.get(k, 0.0) - want.get(k, 0.0)) <= 1e-10, k


from cases import EVAL_CASES as CASES  # noqa: E402


@pytest.mark.parametrize("name", sorted(CASES))
def test_eval_matches_oracle_complex128(name):
    d = CASES[name]()
    want = oracle_eval(d)
    res = d.eval(B200Backend())
    assert res.tensor.array.shape == want.shape                        # exact bookkeeping
    assert res.tensor.array.dtype == np.complex128
    assert _close(res.tensor.array, want, TOL128), name
    if len(res.tensor.dom) == 0 and (res.state_type is StateType.AMP or
                                     any(t in ("bit", "mode") for t in _types(d))):
        if name not in ("bitflip",) and want.ndim > 0:
            _check_prob_dist(res, want, d)


@pytest.mark.parametrize("name", ["lossy_hom", "cfg1_4mode", "ghz6_measured", "clifford_t_8x6",
                                  "cfg3_small"])
def test_eval_matches_oracle_complex64(name):
    d = CASES[name]()
    want = oracle_eval(d)
    res = d.eval(B200Backend(dtype="complex64"))
    assert res.tensor.array.dtype == np.complex64
    assert _close(res.tensor.array, want, TOL64), name


def test_known_answers_through_the_backend():
    p = lossy_hom(0.8).eval(B200Backend()).prob_dist()
    assert p[(1,)] == pytest.approx(0.2, abs=1e-12)                    # README.md:176-178
    r = (photonic.Create(1, 1) >> photonic.BS).eval(B200Backend())
    assert round(r.single_prob((2, 0)), 1) == 0.5                      # backends.py:41-47
    a = (qubits.Z(0, 1) >> qubits.DephasingError(0.43)).eval(B200Backend()).tensor.array
    assert np.allclose(a, np.array([[-.14, 1.], [1, -0.14]]))          # test_qubit.py:112-115
    bell = qubits.Scalar(0.5 ** 0.5) @ qubits.Z(0, 2)
    tele = (qubit @ bell >> qubits.gate("CX") @ qubit >> qubits.H() @ qubit ** 2
            >> qubits.Measure(1) @ qubits.Measure(1) @ qubit >> bit @ classical.CtrlX
            >> classical.CtrlZ)
    t = tele.eval(B200Backend()).tensor.array
    assert t.shape == (2, 2, 2, 2)
    assert np.allclose(t, np.einsum("ac,bd->abcd", np.eye(2), np.eye(2)))   # README.md:88-93
    p = ghz(10, measure=True).eval(B200Backend()).prob_dist()
    assert set(tuple(int(i) for i in k) for k, v in p.items() if v > 1e-12) == \
        {(0,) * 10, (1,) * 10}


def test_sliced_eval_matches_unsliced_and_statevector():
    d = random_clifford_t(10, 8, seed=5)
    want = statevector_clifford_t(10, 8, seed=5)
    full = d.eval(B200Backend()).tensor.array
    backend = B200Backend(target_size=4)
    sliced = d.eval(backend).tensor.array
    assert backend.last_plan.nslices > 1
    assert abs(complex(full) - want) <= 1e-10 and abs(complex(sliced) - want) <= 1e-10


def test_reference_literal_evalresult_cases_on_device_kernels():
    """test/test_backends.py:276-294, 331-345 through the product EvalResult
    (host array uploaded, probs computed by the CUDA kernels)."""
    dm = np.zeros((2, 2, 2), dtype=complex)
    dm[0, 0, 0] = 0.499999999999 + 1e-14j
    dm[0, 1, 1] = 0.100000000001 - 1e-14j
    dm[1, 0, 0] = 0.2000000000005 + 1e-14j
    dm[1, 1, 1] = 0.2000000000005 - 1e-14j
    dm[0, 0, 1] = -1e-13
    dm[1, 1, 0] = -1e-13
    r = EvalResult(ResultTensor((), dm.shape, dm), ("mode", "qubit"), StateType.DM)
    probs = r.prob_dist()
    assert probs[(0,)] == pytest.approx(0.6, rel=1e-12)
    assert probs[(1,)] == pytest.approx(0.4, rel=1e-12)
    off = np.zeros((2, 2, 2))
    off[0, 0, 1], off[1, 1, 0] = 0.6, 0.4
    r = EvalResult(ResultTensor((), off.shape, off), ("mode", "qubit"), StateType.DM)
    with pytest.raises(ValueError):
        r.prob_dist()
    v = np.array([1 / np.sqrt(2), 1j / np.sqrt(2)])
    r = EvalResult(ResultTensor((), (2,), v), ("mode",), StateType.AMP)
    assert r.single_prob((0,)) == pytest.approx(0.5, rel=1e-12)
    assert r.single_prob((2,)) == 0.0


def test_ghz_density_matrix_large_output():
    """cfg2-shaped: 12 qubits, 7 open (2^14-entry density tensor), against the
    closed form rho = (|0..0><0..0| + |1..1><1..1| + (2p-1)^n-weighted coherences)."""
    n, k = 12, 7
    d = ghz(n, open_qubits=k)
    res = d.eval(B200Backend())
    arr = res.tensor.array
    assert arr.shape == (2,) * (2 * k)
    want = np.zeros((2,) * (2 * k), dtype=complex)
    want[(0,) * (2 * k)] = 0.5
    want[(1,) * (2 * k)] = 0.5
    assert _close(arr, want, TOL128)                                   # discarded qubits decohere


@pytest.mark.parametrize("name", ["cfg1_4mode", "ghz8_noisy_measured", "cfg3_small"])
def test_cuda_graph_with_parallel_branches_matches_plain_run(name):
    """The captured graph (contraction tree as parallel branches) replays to the
    same tensor as the stream-ordered run, also after new leaf values arrive."""
    import torch
    from optyx_b200 import executor
    d = CASES[name]()
    net = compile_channel(d)
    want = contract_network(net)
    plan = executor.get_plan(net)
    sess = executor.Session(plan)
    sess.set_leaves([net])
    plain = sess.run([net])[0].cpu().numpy()
    assert plan.parallel_safe
    sess.capture(net.scalar)
    sess.out.zero_()
    sess.replay()
    torch.cuda.synchronize()
    got = sess.out[0].cpu().numpy()
    assert _close(plain, want, TOL128) and _close(got, want, TOL128)
    for _ in range(3):
        sess.replay()
    torch.cuda.synchronize()
    assert _close(sess.out[0].cpu().numpy(), want, TOL128)


def test_eval_batch_one_plan_many_parameter_sets():
    """cfg5: structurally identical diagrams -> one batched plan (leading batch
    mode on every tensor), per-eval scalars as a rank-0 leaf."""
    import math
    from cases import fusion_link, fusion_teleport_input
    backend = B200Backend()
    angles = [i * (math.pi / 2) / 7 for i in range(8)]
    ds = [fusion_link((math.cos(t), math.sin(t)), "bell") for t in angles]
    res = backend.eval_batch(ds)
    assert backend.last_plan.batch == 8
    for d, r in zip(ds, res):
        want = oracle_eval(d)
        assert r.tensor.array.shape == want.shape and _close(r.tensor.array, want, TOL128)
    norm = backend.eval_batch([fusion_link((math.cos(t), math.sin(t)), "discard")
                               for t in angles])
    ratios = [(a.tensor.array / b.tensor.array).real for a, b in zip(res, norm)]
    assert ratios[0] == pytest.approx(1.0, abs=1e-10)
    assert all(x > y for x, y in zip(ratios, ratios[1:]))
    phis = [j / 8 for j in range(8)]
    res = backend.eval_batch([fusion_teleport_input(p) for p in phis])
    for p, r in zip(phis, res):
        assert _close(r.tensor.array, oracle_eval(fusion_teleport_input(p)), TOL128)


def test_eval_batch_pooled_host_compile_matches_serial():
    """eval_batch(make, count): host compile on a fork pool (workers never touch
    the device; the BitControlledGate slab comes from the cache warmed on
    make(0)) gives the same tensors as the serial path, bit for bit."""
    from cases import fusion_teleport_input
    backend = B200Backend()
    make = lambda j: fusion_teleport_input(j / 16)          # noqa: E731
    serial = backend.eval_batch([make(j) for j in range(16)])
    pooled = backend.eval_batch(make=make, count=16, workers=4)
    assert len(pooled) == 16
    for a, b in zip(serial, pooled):
        assert np.array_equal(a.tensor.array, b.tensor.array)
        assert a.state_type == b.state_type


def test_complex64_eval_uses_tensor_core_gemm_and_meets_tolerance():
    """complex64 mode: dense pairs go to the tcgen05 split-precision GEMM; the
    result matches the complex128 oracle within the north star's 1e-5."""
    from cases import cfg1_interferometer
    d = cfg1_interferometer(5, (1, 1, 1, 0, 0), depth=5)
    backend = B200Backend(dtype="complex64")
    res = d.eval(backend)
    want = oracle_eval(d)
    assert res.tensor.array.dtype == np.complex64
    assert np.abs(res.tensor.array - want).max() <= 1e-5 * max(1.0, np.abs(want).max())
    kinds = {st.kernel for st in backend.last_plan.steps}
    assert "gemm" in kinds, kinds


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("name", ["cfg1_4mode", "ghz8_noisy_measured", "fusion_ii", "clifford_t_8x6"])
def test_small_step_program_matches_oracle(name, dtype):
    """Every early tiny step of the plan in ONE persistent launch (b2_run_small_program:
    tasks on shared memory, ticket order, done-flags): same tensor as the oracle and as
    the same plan executed one launch per step; replays (CUDA graph) need no reset."""
    from cases import EVAL_CASES, compile_channel
    from optyx_b200 import executor, planner
    d = EVAL_CASES[name]()
    net = compile_channel(d)
    plan = planner.plan_network(net, dtype=dtype)
    assert plan.small is not None and len(plan.small.step_ids) >= planner.SMALL_MIN_STEPS
    sess = executor.Session(plan)
    got = sess.run([net])[0].cpu().numpy()
    want = oracle_eval(d)
    tol = TOL128 if dtype == np.complex128 else 1e-5
    assert _close(got, want, tol)
    plain = planner.plan_network(net, dtype=dtype)
    for k in plain.small.step_ids:
        plain.steps[k].grouped = False
    plain.small = None
    ref = executor.Session(plain).run([net])[0].cpu().numpy()
    assert _close(got, ref, tol)
    for _ in range(3):                       # plain relaunches: epoch flags, rewound tickets
        sess.out.zero_()
        sess.build_leaves()
        sess.contract(net.scalar)
        assert _close(sess.out[0].cpu().numpy(), want, tol)
    sess.capture(net.scalar)                 # and as a node of the captured graph
    for _ in range(3):
        sess.out.zero_()
        sess.replay()
        assert _close(sess.out[0].cpu().numpy(), want, tol)


def test_small_step_program_batched():
    """A batch: (eval, task) work items, per-eval strides on imports / exports."""
    from cases import compile_channel, fusion_teleport_input
    from optyx_b200 import executor
    from optyx_b200.core import network as nw
    nets = [compile_channel(fusion_teleport_input(0.01 * j)) for j in range(40)]
    wants = [contract_network(n) for n in nets]
    for net in nets:
        net.leaves.append(nw.Leaf(nw.K_DENSE, (), (), 0, np.array([net.scalar], dtype=complex), "scalar"))
        net.scalar = 1.0 + 0j
    plan = executor.get_plan(nets[0], batch=len(nets))
    assert plan.small is not None
    sess = executor.Session(plan)
    out = sess.run(nets, scalar=1.0 + 0j).cpu().numpy()
    for e, want in enumerate(wants):
        assert _close(out[e], want, TOL128), e


def test_cfg2_full_size_closed_form_on_device():
    """BASELINE cfg2 at the bench size: 20-qubit GHZ, DephasingError(0.95) on every
    qubit, 14 open / 6 discarded -> 2^28-entry density tensor (4.3 GB, checked on
    the device).  Closed form: the discarded qubits kill the coherences, and by the
    reference's zx.ZBox.conjugate quirk each noise box multiplies
    the diagonal by (2p - 1): exactly two non-zero entries 0.5 * 0.9^20."""
    import torch
    from cases import compile_channel
    from optyx_b200 import executor
    n, k, p = 20, 14, 0.95
    net = compile_channel(ghz(n, open_qubits=k, noise=p))
    dev = executor.evaluate(net)
    assert tuple(dev.shape) == (2,) * (2 * k)
    flat = dev.reshape(-1)
    want = 0.5 * (2 * p - 1) ** n
    first, last = complex(flat[0].item()), complex(flat[-1].item())
    assert abs(first - want) <= 1e-10 * want and abs(last - want) <= 1e-10 * want
    mag = flat.abs()
    assert int((mag > 1e-12 * want).sum().item()) == 2          # nothing else survives
    assert abs(float(mag.sum().item()) - 2 * want) <= 1e-9 * want
    del dev, flat, mag
    torch.cuda.empty_cache()


def test_cfg4_full_size_two_slicings_agree():
    """BASELINE cfg4 at the bench size (30 qubits, depth 41): the amplitude summed
    over 2^26-element slices equals the one summed over 2^25-element slices (two
    different plans, different slice sets), and |amplitude| <= 1."""
    from cases import compile_channel
    from optyx_b200 import executor
    net = compile_channel(random_clifford_t(30, 41, seed=0))
    vals = []
    for tgt in (26, 25):
        plan = executor.get_plan(net, target_size=float(2 ** tgt), trials=8)
        assert plan.nslices > 1 and plan.info.peak <= 2 ** tgt
        sess = executor.Session(plan)
        vals.append(complex(sess.run([net])[0].cpu().numpy()))
        del sess
    a, b = vals
    assert abs(a) <= 1.0 and abs(a - b) <= 1e-9 * abs(a), (a, b)


def test_formal_sums_on_device():
    """backends.py:473-476: a Sum is evaluated term by term and added (b2_axpy);
    pure sum = test_zw.py:284-292 (branching), mixed sum: two loss channels."""
    from optyx_b200 import photonic
    from optyx_b200.core import zw
    branching = photonic.Create(1) @ photonic.Create(0) + photonic.Create(0) @ photonic.Create(1)
    res = branching.eval(B200Backend())
    assert res.state_type is StateType.AMP
    assert _close(res.tensor.array, oracle_eval(branching), TOL128)
    amps = res.amplitudes()
    assert set(amps) == {(1, 0), (0, 1)} and abs(amps[(1, 0)] - 2 ** -0.5) < 1e-12
    mixed = photonic.Create(1) >> (photonic.PhotonLoss(0.3) + photonic.PhotonLoss(0.9)) \
        >> photonic.NumberResolvingMeasurement(1)
    r = mixed.eval(B200Backend())
    assert r.state_type is StateType.DM and _close(r.tensor.array, oracle_eval(mixed), TOL128)
    p = r.prob_dist()
    assert p[(1,)] == pytest.approx(0.6, abs=1e-12) and p[(0,)] == pytest.approx(0.4, abs=1e-12)


@pytest.mark.parametrize("name", ["BS", "phase_tbs", "MZI", "chip", "bunched"])
def test_permanent_backend_matches_tensor_network_backend(name):
    """test_backends.py:482-509: PermanentBackend == QuimbBackend on Create >> circuit
    (amplitudes, prob_dist, tensor to 1e-12) -- here GPU permanents vs the CUDA
    tensor-network path, and both vs the oracle."""
    from optyx_b200 import photonic
    from optyx_b200.core.backends import PermanentBackend
    circ, photons = {
        "BS": (photonic.BS, (1, 1)),
        "phase_tbs": (photonic.Phase(0.2) @ photonic.Phase(0.3) >> photonic.TBS(0.3), (1, 1)),
        "MZI": (photonic.MZI(0.2, 0.8), (1, 1)),
        "chip": (photonic.seeded_ansatz(4, 4, seed=7), (1, 1, 1, 1)),
        "bunched": (photonic.seeded_ansatz(3, 3, seed=2), (2, 0, 1)),
    }[name]
    d = photonic.Create(*photons) >> circ
    tn = d.eval(B200Backend())
    pm = d.eval(PermanentBackend())
    assert pm.state_type is StateType.AMP
    n = sum(photons)
    want = oracle_eval(d)
    # the tensor-network result is truncated by the light cone; embed it in (n+1)^m
    full = np.zeros((n + 1,) * len(photons), dtype=complex)
    full[tuple(slice(0, s) for s in want.shape)] = want
    assert np.abs(pm.tensor.array - full).max() <= 1e-12
    a_tn, a_pm = tn.amplitudes(), pm.amplitudes()
    assert set(a_tn) == set(a_pm)
    assert all(abs(a_tn[k] - a_pm[k]) <= 1e-10 for k in a_tn)
    p_tn, p_pm = tn.prob_dist(), PermanentBackend().prob_dist(d)
    assert all(abs(p_tn[k] - p_pm.get(k, 0.0)) <= 1e-10 for k in p_tn)


def test_permanent_backend_12_modes_6_photons_is_normalised():
    """The BASELINE cfg3 interferometer without loss and distinguishability: 12 modes,
    6 photons, 12 376 output configurations, far beyond an exact tensor-network
    contraction; unitarity => the probabilities sum to 1."""
    from optyx_b200 import photonic
    from optyx_b200.core.backends import PermanentBackend
    d = photonic.Create(1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0) >> photonic.seeded_ansatz(12, 12, seed=0)
    p = PermanentBackend().prob_dist(d)
    assert len(p) <= 12376 and all(sum(k) == 6 for k in p)
    assert sum(p.values()) == pytest.approx(1.0, abs=1e-10)
    with pytest.raises(AssertionError):
        (d >> photonic.NumberResolvingMeasurement(12)).eval(PermanentBackend())


@pytest.mark.parametrize("shape", [(4, 2, 2), (5, 3, 2), (4, 3, 4)])
def test_lossy_distinguishable_permanents_match_tensor_network(shape):
    """cfg3-shaped circuits (loss on every mode, partially distinguishable photons,
    some modes measured, the rest discarded): detected-count distribution from GPU
    permanents (device enumeration + marginalisation) == CP-doubled tensor network."""
    from cases import cfg3_interferometer
    from optyx_b200.core.backends import PermanentBackend
    width, n_photons, measured = shape
    tn = cfg3_interferometer(width=width, n_photons=n_photons, measured=measured).eval(
        B200Backend()).prob_dist()
    pm = PermanentBackend().prob_dist_measured(
        cfg3_interferometer(width=width, n_photons=n_photons, measured=measured, inflate=False))
    assert sum(pm.values()) == pytest.approx(1.0, abs=1e-10)
    keys = set(k for k, v in tn.items() if v > 1e-13) | set(k for k, v in pm.items() if v > 1e-13)
    for k in keys:
        assert abs(tn.get(tuple(k), 0.0) - pm.get(tuple(k), 0.0)) <= 1e-10, (k, tn.get(k), pm.get(k))


def test_lossy_permanents_beyond_tensor_network_reach_match_oracle_permanents():
    """8 modes, 4 photons, loss, partial distinguishability, 4 measured (52 360
    configurations over 32 modes; the tensor network would need 2^55 MACs): device
    histogram == the same sum done in Python with the ORACLE's permanent."""
    import itertools
    from math import factorial
    from cases import cfg3_interferometer
    from oracle.permanent import npperm
    from optyx_b200.core.backends import PermanentBackend
    d = cfg3_interferometer(width=8, n_photons=4, measured=4, inflate=False)
    got = PermanentBackend().prob_dist_measured(d)
    V, where, k = PermanentBackend._photon_rows(d)
    n, M = V.shape
    want = {}
    for cols in itertools.combinations_with_replacement(range(M), n):
        key = [0] * k
        for c in cols:
            if where[c] >= 0:
                key[where[c]] += 1
        occ = np.bincount(cols, minlength=M)
        w = abs(npperm(V[:, list(cols)])) ** 2 / np.prod([factorial(int(x)) for x in occ])
        want[tuple(key)] = want.get(tuple(key), 0.0) + w
    assert set(k_ for k_, v in want.items() if v > 1e-14) <= set(got)
    assert max(abs(got.get(k_, 0.0) - v) for k_, v in want.items()) <= 1e-12


def test_cfg3_full_size_through_permanents():
    """BASELINE cfg3 at its full size -- 12 modes, 6 photons, PhotonLoss(0.9) on every
    mode, partially distinguishable photons (internal dimension 2), 4 modes measured,
    8 discarded: 22 957 480 output configurations over 48 modes, beyond any exact
    tensor-network contraction.  Size-independent properties:
    normalisation, photon-number bound, loss monotonicity of the vacuum outcome."""
    from cases import cfg3_interferometer
    from optyx_b200.core.backends import PermanentBackend
    pb = PermanentBackend()
    p = pb.prob_dist_measured(cfg3_interferometer(width=12, n_photons=6, measured=4, inflate=False))
    assert pb.last_configs == 22957480
    assert sum(p.values()) == pytest.approx(1.0, abs=1e-9)
    assert all(len(k) == 4 and sum(k) <= 6 for k in p) and all(v >= 0 for v in p.values())
    q = PermanentBackend().prob_dist_measured(
        cfg3_interferometer(width=12, n_photons=6, measured=4, loss=0.5, inflate=False))
    assert q[(0, 0, 0, 0)] > p[(0, 0, 0, 0)]          # more loss -> more often nothing detected


def test_edge_case_diagrams():
    """Degenerate networks: identity (no leaves), scalar only, a single box,
    disconnected components (outer product), an identically zero result."""
    from optyx_b200 import photonic, qubits
    from optyx_b200.core import zw
    cases_ = {
        "identity": qubits.Id(2),
        "scalar": qubits.Scalar(0.25 + 0.5j),
        "single_box": qubits.H(),
        "outer_product": qubits.Ket(0) @ qubits.Ket(1) @ qubits.H(),
        "zero": zw.Create(1) >> zw.Select(0),                  # test_zw.py:275-280 (bone)
        "two_components": photonic.Create(1) @ photonic.Create(2),
    }
    for name, d in cases_.items():
        want = oracle_eval(d)
        got = d.eval(B200Backend()).tensor.array if hasattr(d, "eval") else None
        if got is None:
            from optyx_b200 import executor
            got = executor.evaluate(d.to_network()).cpu().numpy()
        assert np.shape(got) == np.shape(want), name
        assert _close(np.asarray(got), np.asarray(want), TOL128), name
    # an exactly-zero tensor: prob_dist raises like the reference (backends.py:231-232)
    r = EvalResult(ResultTensor((), (3,), np.zeros(3, dtype=complex)), ("mode",), StateType.AMP)
    with pytest.raises(ValueError):
        r.prob_dist()


def test_quimb_adapter_path_on_device():
    """INTEGRATION.md section 4, adapter path: dense tensors with index names (what
    ``term.to_quimb()`` yields) contracted by the CUDA kernels."""
    from types import SimpleNamespace
    from optyx_b200 import executor
    from optyx_b200.adapters import network_from_quimb
    rng = np.random.default_rng(11)
    shapes = {"a": 3, "b": 4, "c": 2, "d": 5, "e": 3, "h": 4}
    specs = [("a", "b", "h"), ("h", "c", "d"), ("d", "e", "h"), ("b", "c")]
    tensors = [SimpleNamespace(data=rng.standard_normal([shapes[i] for i in ix]) +
                               1j * rng.standard_normal([shapes[i] for i in ix]), inds=ix)
               for ix in specs]
    tn = SimpleNamespace(tensors=tensors, outer_inds=lambda: ("a", "e"))
    net = network_from_quimb(tn)
    got = executor.evaluate(net).cpu().numpy()
    want = np.einsum("abh,hcd,deh,bc->ae", *[t.data for t in tensors])
    assert _close(got, want, TOL128)


def test_cfg3_standard_size_properties():
    """cfg3 at a size the oracle cannot reach quickly: 6 modes x inflate 2,
    2 partially distinguishable photons, loss, 3 measured modes.  Checked through
    size-independent properties: normalisation, non-negativity, photon number."""
    from cases import cfg3_interferometer
    d = cfg3_interferometer(width=6, n_photons=2, measured=3)
    res = d.eval(B200Backend())
    p = res.prob_dist()
    assert sum(p.values()) == pytest.approx(1.0, abs=1e-9)
    assert all(v >= -1e-12 for v in p.values())
    assert all(sum(int(i) for i in k) <= 2 for k, v in p.items() if v > 1e-12)


def test_sparse_result_hand_over_equals_dense_copy(monkeypatch):
    """B200Backend._to_host: a large, mostly-zero result crosses PCIe as its list of
    non-zero entries (b2_compact_nonzero) and equals the dense copy; a dense result,
    or one below the size threshold, takes the dense copy."""
    import torch
    rng = np.random.default_rng(5)
    for dtype, tdtype in ((np.complex128, torch.complex128), (np.complex64, torch.complex64)):
        host = np.zeros((2,) * 22, dtype=dtype)                # 64 MB / 32 MB
        flat = host.reshape(-1)
        where = rng.choice(flat.size, size=1000, replace=False)
        flat[where] = (rng.standard_normal(1000) + 1j * rng.standard_normal(1000)).astype(dtype)
        dev = torch.from_numpy(host).cuda()
        got = B200Backend(dtype=np.dtype(dtype).name)._to_host(dev)
        assert B200Backend.last_d2h["mode"] == "sparse" and B200Backend.last_d2h["nonzeros"] == 1000
        assert B200Backend.last_d2h["bytes"] == 8 + 1000 * (8 + host.itemsize)
        assert got.shape == host.shape and got.dtype == host.dtype and got.flags.c_contiguous
        assert np.array_equal(got, host)
        flat[::3] = 1.0                                        # a third non-zero: dense copy
        dev = torch.from_numpy(host).cuda()
        got = B200Backend(dtype=np.dtype(dtype).name)._to_host(dev)
        assert B200Backend.last_d2h["mode"] == "dense"
        assert np.array_equal(got, host)
        got = B200Backend(sparse_d2h=False)._to_host(dev)
        assert B200Backend.last_d2h == {"mode": "dense", "bytes": host.nbytes}
        del dev
    # through eval: both hand-overs return the same array
    d = ghz(20, open_qubits=11, noise=0.95)                   # (2,)^22: 64 MB
    a = d.eval(B200Backend()).tensor.array
    b = d.eval(B200Backend(sparse_d2h=False)).tensor.array
    assert a.shape == b.shape == (2,) * 22 and _close(a, b, 1e-13)
    ghz(8, open_qubits=4).eval(B200Backend())
    assert B200Backend.last_d2h["mode"] == "dense"            # below the size threshold
    monkeypatch.setattr(B200Backend, "SPARSE_MIN_BYTES", 0)   # scan small results too
    modes = {}
    for name in ("cfg1_4mode", "cfg1_6mode", "ghz8_noisy_measured", "mzi_chip_pure",
                 "clifford_t_8x6", "fusion_ii"):
        a = CASES[name]().eval(B200Backend()).tensor.array
        modes[name] = B200Backend.last_d2h["mode"]
        b = CASES[name]().eval(B200Backend(sparse_d2h=False)).tensor.array
        # two evals differ in the last bits where split-K atomics order the sums
        assert a.shape == b.shape and _close(a, b, 1e-13), (name, modes)


@pytest.mark.parametrize("name", ["per_mode_discards", "discard_id_discard", "discard_then_gate",
                                  "interleaved", "swap_after_discard"])
def test_permanent_histogram_with_discards_before_later_boxes(name):
    """a Discard in the middle of the circuit must not shift later boxes onto
    the wrong modes -- device histogram vs the CP-doubled oracle."""
    from test_planner_cpu import _discard_cases
    from optyx_b200.core.backends import PermanentBackend
    d = _discard_cases()[name]
    got = PermanentBackend().prob_dist_measured(d)
    want = ev.prob_dist_mixed(oracle_eval(d), list(d.cod.inside))
    total = sum(want.values())
    keys = set(want) | set(got)
    assert max(abs(got.get(k, 0.0) - want.get(k, 0.0) / total) for k in keys) <= 1e-10


def test_eval_params_matches_per_diagram_evals_and_oracle():
    """One compile with array-valued parameters (eval_params) == eval_batch == oracle."""
    import math
    from cases import fusion_link, fusion_teleport_input
    backend = B200Backend()
    phis = np.arange(24) / 24
    got = backend.eval_params(fusion_teleport_input, phis)
    ref = backend.eval_batch(make=lambda j: fusion_teleport_input(phis[j]), count=24, workers=1)
    assert len(got) == 24
    for j in (0, 5, 23):
        want = oracle_eval(fusion_teleport_input(phis[j]))
        assert _close(got[j].tensor.array, want, TOL128)
    for a, b in zip(got, ref):
        assert _close(a.tensor.array, b.tensor.array, 1e-12)
    th = np.linspace(0, math.pi / 2, 16)
    got = backend.eval_params(lambda t: fusion_link((np.cos(t), np.sin(t)), "bell"), th)
    for j in (0, 7, 15):
        want = oracle_eval(fusion_link((math.cos(th[j]), math.sin(th[j])), "bell"))
        assert _close(got[j].tensor.array, want, TOL128)
    # second call: cached session + captured graph, new parameters
    got2 = backend.eval_params(fusion_teleport_input, phis[::-1].copy())
    assert _close(got2[0].tensor.array, oracle_eval(fusion_teleport_input(phis[-1])), TOL128)
