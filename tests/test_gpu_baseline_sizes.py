"""-m gpu: every BENCH workload at its BASELINE size against a checker that is NOT the
CUDA path itself ("parity at BASELINE sizes"):

* cfg4 (30 qubits, the bench circuit) vs an independent gate-by-gate statevector
  simulation (torch ops on a 17 GB state: test-only, shares nothing with the product);
* cfg2 (the 2^28-entry density tensor of the bench) element by element vs the CPU
  oracle (independent host compile + numpy contraction);
* cfg5a / cfg5b: all 1024 evals of the batch vs the CPU oracle;
* cfg3 (the tensor-network instance of the bench) vs the ORACLE's permanents;
* cfg1 (bench instance) vs the CPU oracle.
"""
import math

import numpy as np
import pytest

import bench
from cases import oracle_eval, statevector_clifford_t, statevector_clifford_t_torch
from oracle import evalresult as ev
from optyx_b200.core.backends import B200Backend, PermanentBackend

pytestmark = pytest.mark.gpu


def test_statevector_checker_agrees_with_the_numpy_one():
    """The torch checker itself, against the numpy statevector of tests/cases.py."""
    for n, depth, seed in ((6, 5, 1), (9, 7, 2)):
        a = statevector_clifford_t_torch(n, depth, seed)
        b = complex(statevector_clifford_t(n, depth, seed))
        assert abs(a - b) <= 1e-12


def test_cfg4_bench_circuit_matches_an_independent_statevector():
    """The bench's cfg4 instance (30 qubits, same depth / seed / slicing as bench.py)
    against <0|C|0> from a full 2^30 statevector; relative error <= 1e-10."""
    import torch
    w = bench.workload("cfg4")
    d = w["factory"]()
    backend = B200Backend(target_size=w["target"], trials=w["trials"], min_slices=w["min_slices"])
    got = complex(np.asarray(d.eval(backend).tensor.array))
    plan = backend.last_plan
    assert plan.nslices >= bench.CFG4_MIN_SLICES and plan.info.peak <= bench.CFG4_TARGET
    del backend
    torch.cuda.empty_cache()
    n, depth = 30, int(w["desc"].split("depth ")[1].split()[0])
    want = statevector_clifford_t_torch(n, depth, seed=0)
    assert abs(want) > 0
    assert abs(got - want) <= 1e-10 * abs(want), (got, want)


def test_cfg2_bench_tensor_elementwise_against_the_cpu_oracle():
    w = bench.workload("cfg2")
    d = w["factory"]()
    want = oracle_eval(d)                               # ~5 s of numpy on the host
    got = d.eval(B200Backend()).tensor.array
    assert got.shape == want.shape == (2,) * 28
    err = float(np.abs(got - want).max())
    assert err <= 1e-10 * float(np.abs(want).max()), err


def _oracle_batch(make, count):
    """oracle_eval of every diagram of a batch, on a fork pool (pure numpy)."""
    import multiprocessing as mp
    import os
    global _MAKE
    _MAKE = make
    with mp.get_context("fork").Pool(min(16, os.cpu_count() or 1)) as pool:
        return pool.map(_oracle_one, range(count), chunksize=16)


def _oracle_one(j):
    return oracle_eval(_MAKE(j))


@pytest.mark.parametrize("name", ["cfg5a", "cfg5b"])
def test_cfg5_all_1024_evals_against_the_cpu_oracle(name):
    w = bench.workload(name)
    res = B200Backend().eval_batch(make=w["make"], count=w["batch"])
    assert len(res) == 1024
    wants = _oracle_batch(w["make"], w["batch"])
    scale = max(float(np.abs(x).max()) for x in wants)
    worst = 0.0
    for r, want in zip(res, wants):
        got = np.asarray(r.tensor.array)
        assert got.shape == np.shape(want)
        worst = max(worst, float(np.abs(got - want).max()))
    assert worst <= 1e-10 * max(1.0, scale), worst


def test_cfg3_bench_instance_against_oracle_permanents():
    """6 modes, 3 photons, loss, partial distinguishability, inflate(2), 2 modes
    measured (the 2^39-MAC instance bench.py --workload cfg3 times): the tensor-network
    prob_dist vs the detected-count distribution summed in Python from the ORACLE's
    permanents (path.py:141-163) over all C(26, 3) output configurations."""
    import itertools
    from math import factorial
    from cases import cfg3_interferometer
    from oracle.permanent import npperm
    d = cfg3_interferometer(width=6, n_photons=3, measured=2)
    got = d.eval(B200Backend()).prob_dist()
    V, where, k = PermanentBackend._photon_rows(
        cfg3_interferometer(width=6, n_photons=3, measured=2, inflate=False))
    n, M = V.shape
    want = {}
    for cols in itertools.combinations_with_replacement(range(M), n):
        key = [0] * k
        for c in cols:
            if where[c] >= 0:
                key[where[c]] += 1
        occ = np.bincount(cols, minlength=M)
        wgt = abs(npperm(V[:, list(cols)])) ** 2 / np.prod([factorial(int(x)) for x in occ])
        want[tuple(key)] = want.get(tuple(key), 0.0) + wgt
    assert sum(want.values()) == pytest.approx(1.0, abs=1e-12)
    for key in set(want) | set(got):
        assert abs(got.get(key, 0.0) - want.get(key, 0.0)) <= 1e-10, key


def test_cfg1_bench_instance_against_the_cpu_oracle():
    w = bench.workload("cfg1")
    d = w["factory"]()
    want = oracle_eval(d)
    res = d.eval(B200Backend())
    assert res.tensor.array.shape == want.shape == (4,) * 6
    assert np.abs(res.tensor.array - want).max() <= 1e-10 * max(1.0, np.abs(want).max())
    wp = ev.prob_dist(want, list(d.cod.inside), ev.DM)
    gp = res.prob_dist()
    for key in set(wp) | set(gp):
        assert abs(gp.get(key, 0.0) - wp.get(tuple(int(i) for i in key), 0.0)) <= 1e-10
