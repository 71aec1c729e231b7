"""ORACLE (test infrastructure, NOT product code) -- result post-processing.

Literal CPU restatement of ``EvalResult`` (optyx/core/backends.py:128-356):
``_convert_array_to_dict`` 274-289, ``_prob_dist_pure`` 291-304,
``_prob_dist_mixed`` 306-356, ``prob_dist`` 196-240, ``amplitudes`` 169-194.
Pinned by the reference's own literal-array tests
(test/test_backends.py:255-405, 456-465), reproduced in
``tests/test_oracle_golden.py``.

``output_types`` is a sequence of strings ("bit", "mode", "qubit", "qmode");
anything that is not "bit"/"mode" counts as an unmeasured quantum wire, as in
the reference's ``t in {bit, mode}`` test (backends.py:332-334).
"""
from __future__ import annotations

from collections import defaultdict

import numpy as np

AMP, DM, PROB = "amp", "dm", "prob"


def convert_array_to_dict(array: np.ndarray) -> dict:
    # backends.py:274-289
    array = np.asarray(array)
    nz_flat = np.flatnonzero(array)
    if nz_flat.size == 0:
        return {}
    nz_vals = array.flat[nz_flat]
    nz_multi = np.vstack(np.unravel_index(nz_flat, array.shape)).T
    return {tuple(int(i) for i in idx): val for idx, val in zip(nz_multi, nz_vals)}


def prob_dist_pure(array) -> dict:
    # backends.py:291-304
    return {k: abs(v) ** 2 for k, v in convert_array_to_dict(array).items()}


def prob_dist_mixed(array, output_types) -> dict:
    # backends.py:306-356
    if not any(t in {"bit", "mode"} for t in output_types):
        raise ValueError("Output types must contain at least one 'bit' or 'mode'.")
    values = convert_array_to_dict(array)
    mask_flat = np.concatenate(
        [[1] if t in {"bit", "mode"} else [0, 0] for t in output_types])
    probs = defaultdict(float)
    all_measured = set()
    for key, amp in values.items():
        occ_measured = tuple(i for i, m in zip(key, mask_flat) if m)
        all_measured.add(occ_measured)
        occs_unmeasured = tuple(i for i, m in zip(key, mask_flat) if not m)
        if all(occs_unmeasured[i] == occs_unmeasured[i + 1]
               for i in range(0, len(occs_unmeasured) - 1, 2)):
            val = float(np.real_if_close(amp))
            if val < 0 and abs(val) < 1e-12:
                val = 0.0
            probs[occ_measured] += val
    for occ in all_measured:
        probs.setdefault(occ, 0.0)
    return dict(probs)


def prob_dist(array, output_types, state_type, n_dom: int = 0, round_digits=None) -> dict:
    # backends.py:196-240
    if n_dom != 0:
        raise ValueError("Result tensor must represent a state with no inputs.")
    if state_type == AMP:
        values = prob_dist_pure(array)
    elif state_type == DM:
        values = prob_dist_mixed(array, output_types)
    elif state_type == PROB:
        values = convert_array_to_dict(array)
    else:
        raise ValueError(f"Unsupported state_type type: {state_type}.")
    if not np.allclose(np.sum(list(values.values())), 1, atol=1e-12):
        total = float(np.sum(list(values.values())))
        if total == 0:
            raise ValueError("The probability distribution sums to zero.")
        prob = {k: v / total for k, v in values.items()}
    else:
        prob = values
    if round_digits is None:
        return prob
    return {k: round(v, round_digits) for k, v in prob.items()}


def amplitudes(array, state_type, n_dom: int = 0, normalise=True) -> dict:
    # backends.py:169-194
    if state_type != AMP:
        raise TypeError(f"Cannot get amplitudes from {state_type}.")
    if n_dom != 0:
        raise ValueError("Result tensor must represent a state with no inputs.")
    dic = convert_array_to_dict(array)
    if normalise:
        norm = np.sqrt(np.sum(np.abs(list(dic.values())) ** 2))
        return {key: value / norm for key, value in dic.items()}
    return dic
