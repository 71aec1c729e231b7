"""Regenerates the fixtures in this directory.

The reference itself cannot be imported here (discopy / quimb / cotengra / perceval
are absent), so there are two kinds of fixture, both small:

* ``reference_known_answers.json`` -- literal numbers transcribed from the reference's
  own tests and docstrings (each entry cites its file:line).  Written by hand in this
  script, NOT computed: they are what pins the oracle.
* ``eval_cases.npz`` -- the result tensors of the named eval cases of ``tests/cases.py``
  (BASELINE-shaped small circuits) as produced by host compile + the numpy oracle
  (``oracle/``), stored so that the CUDA path on the GPU box is also held to committed
  vectors, and so that a change of the oracle or the host compiler shows up as a diff.

usage:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402

KNOWN = {
    "hom_amplitudes": {
        "cite": "test/test_zw.py:410-442",
        "what": "Create(1)@Create(1) >> BS(ZW) >> Select(s1)@Select(s2)",
        "values": {"1,1": [0.0, 0.0], "2,0": [0.0, 0.7071067811865476],
                   "0,2": [0.0, 0.7071067811865476]},
    },
    "lossy_hom_prob": {
        "cite": "README.md:176-178",
        "what": "Create(1,1) >> PhotonLoss(0.8)@Id >> BS >> NumberResolvingMeasurement(2) >> AddN(2)",
        "values": {"1": 0.2, "2": 0.8},
    },
    "bs_single_prob": {
        "cite": "optyx/core/backends.py:41-47",
        "what": "(Create(1,1) >> BS).eval().single_prob((2, 0)) rounded to 1 digit",
        "values": {"2,0": 0.5},
    },
    "bit_flip_error": {
        "cite": "test/test_qubit.py:107-110",
        "what": "(Z(0,1) >> BitFlipError(0.43)).double() array",
        "values": [[-0.14, -0.14], [-0.14, -0.14]],
    },
    "dephasing_error": {
        "cite": "test/test_qubit.py:112-115",
        "what": "(Z(0,1) >> DephasingError(0.43)).double() array",
        "values": [[-0.14, 1.0], [1.0, -0.14]],
    },
    "noisy_bell_fidelity": {
        "cite": "test/test_channel.py:22-34",
        "what": "fidelity of the depolarised Bell pair",
        "values": 0.96898,
    },
    "mixed_prob_dist_literal": {
        "cite": "test/test_backends.py:404-405",
        "what": "EvalResult(DM literal).prob_dist()",
        "values": {"0": 0.55, "1": 0.45},
    },
}


def main():
    with open(os.path.join(HERE, "reference_known_answers.json"), "w") as fh:
        json.dump(KNOWN, fh, indent=1, sort_keys=True)
        fh.write("\n")
    arrays = {}
    for name in sorted(cases.EVAL_CASES):
        arr = np.asarray(cases.oracle_eval(cases.EVAL_CASES[name]()), dtype=np.complex128)
        arrays[name] = arr
        print(f"{name:24s} shape {arr.shape}  |max| {np.abs(arr).max():.6g}")
    np.savez_compressed(os.path.join(HERE, "eval_cases.npz"), **arrays)


if __name__ == "__main__":
    main()
