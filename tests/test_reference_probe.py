"""Run-time probe for the REAL reference.

The reference is Python but needs discopy, quimb, cotengra, perceval, pyzx, graphix
and pytket at import; none is in this image (``python -c "import discopy"`` fails,
``/opt/wheelhouse`` holds no wheel of them, there is no network), so every test in
this file SKIPS here and on the GPU box.  The moment a reference install exists --
``baseline/_ref/`` (the driver's reference-arm directory) or site-packages -- they
activate and compare, on the same circuits,

* the independent oracle (``oracle/compile.py``) with ``QuimbBackend().eval``,
* the adapter ``network_from_quimb(term.to_quimb())`` (the replacement for
  backends.py:506) + the CUDA path with ``tn.contract(output_inds=...)``,
* ``B200Backend`` on our mirrored diagram with ``QuimbBackend`` on the reference's.

The reference circuits are built through the reference's own public API
(README.md:50-93, 170-178; photonic.py:350-361, 1140-1145; qubits.py:575-592).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _probe():
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref_dir) and ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    try:
        import cotengra  # noqa: F401
        import discopy  # noqa: F401
        import optyx  # noqa: F401
        import quimb  # noqa: F401
    except Exception as exc:           # ImportError, or a half-installed dependency failing
        return None, f"{type(exc).__name__}: {exc}"
    return optyx, ""


REF, WHY = _probe()
needs_ref = pytest.mark.skipif(REF is None, reason=f"the reference does not import here ({WHY})")


def probe_status() -> dict:
    """What bench.py reports about the reference arm."""
    return {"importable": REF is not None, "why": WHY}


def _pairs():
    """(name, reference diagram, mirrored optyx_b200 diagram) for circuits both APIs spell
    the same way."""
    import optyx.classical as rc
    import optyx.photonic as rp
    import optyx.qubits as rq
    from optyx_b200 import classical as c, photonic as p, qubits as q

    def both(fn):
        return fn(rp, rq, rc), fn(p, q, c)

    cases = {
        "hom": lambda P, Q, C: P.Create(1, 1) >> P.BS,
        "lossy_hom": lambda P, Q, C: (P.Create(1, 1) >> P.PhotonLoss(0.8) @ P.Id(1) >> P.BS
                                      >> P.NumberResolvingMeasurement(2) >> C.AddN(2)),
        "mzi_tbs": lambda P, Q, C: (P.Create(1, 1, 0) >> P.MZI(0.3, 0.7) @ P.Id(1)
                                    >> P.Id(1) @ P.TBS(0.2) >> P.Phase(0.1) @ P.Id(2)),
        "bell_dephased": lambda P, Q, C: (Q.Ket(0) @ Q.Ket(0) >> Q.H() @ Q.Id(1)
                                          >> Q.Z(1, 2) @ Q.Id(1) >> Q.Id(1) @ Q.X(2, 1)
                                          >> Q.DephasingError(0.9) @ Q.BitFlipError(0.8)),
        "bell_measured": lambda P, Q, C: (Q.Ket(0) @ Q.Ket(0) >> Q.H() @ Q.Id(1)
                                          >> Q.Z(1, 2) @ Q.Id(1) >> Q.Id(1) @ Q.X(2, 1)
                                          >> Q.Measure(2)),
        "dual_rail_threshold": lambda P, Q, C: (Q.Ket("+") >> P.DualRail(1)
                                                >> P.PhotonThresholdMeasurement(2)),
    }
    return [(name,) + both(fn) for name, fn in cases.items()]


@needs_ref
def test_independent_oracle_matches_quimb_backend():
    from optyx.core.backends import QuimbBackend
    from oracle.compile import eval_dense
    for name, ref_d, our_d in _pairs():
        want = np.asarray(QuimbBackend().eval(ref_d).tensor.array)
        got = eval_dense(our_d)
        assert got.shape == want.shape, name
        assert np.abs(got - want).max() <= 1e-10 * max(1.0, np.abs(want).max()), name


@needs_ref
@pytest.mark.gpu
def test_quimb_adapter_and_cuda_path_match_quimb_contract():
    """backends.py:506-545 with our executor in place of quimb's contraction: the
    reference's own tensors (``term.to_quimb()``), explicit ``output_inds``."""
    from optyx.core.backends import QuimbBackend
    from optyx_b200 import executor
    from optyx_b200.adapters import network_from_quimb
    for name, ref_d, _ in _pairs():
        td = QuimbBackend()._get_discopy_tensor(ref_d)
        terms = td.terms if hasattr(td, "terms") else [td]
        for term in terms:
            tn = term.to_quimb()
            out = sorted(tn.outer_inds())
            want = tn.contract(output_inds=out)
            want = np.asarray(getattr(want, "data", want))
            got = executor.evaluate(network_from_quimb(tn, output_inds=out)).cpu().numpy()
            assert np.abs(got - want).max() <= 1e-10 * max(1.0, np.abs(want).max()), name


@needs_ref
@pytest.mark.gpu
def test_b200_backend_matches_quimb_backend():
    from optyx.core.backends import QuimbBackend
    from optyx_b200.core.backends import B200Backend
    for name, ref_d, our_d in _pairs():
        want = QuimbBackend().eval(ref_d)
        got = our_d.eval(B200Backend())
        a, b = np.asarray(got.tensor.array), np.asarray(want.tensor.array)
        assert a.shape == b.shape, name
        assert np.abs(a - b).max() <= 1e-10 * max(1.0, np.abs(b).max()), name
        if len(want.tensor.dom) == 0 and name != "hom":
            pw = want.prob_dist()
            pg = got.prob_dist()
            for k in set(pw) | set(pg):
                assert abs(pw.get(k, 0.0) - pg.get(k, 0.0)) <= 1e-10, (name, k)


def test_probe_reports_why_it_skips():
    st = probe_status()
    assert isinstance(st["importable"], bool)
    assert st["importable"] or st["why"]
