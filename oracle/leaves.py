"""ORACLE (test infrastructure, NOT product code) -- dense leaf arrays.

CPU restatement of how the reference turns a box into a dense truncated
array: ``Box.truncation`` (optyx/core/diagram.py:536-590) drives a per-box
``truncation_specification`` generator over every input basis state and
scatters ``BasisTransition(out, amp)`` (optyx/utils/utils.py:301-320) into
``A[out + inp]``.  Each function below follows one generator literally
(Python loops over basis states), so it is independent of the closed-form
predicates the CUDA array-build kernel evaluates per element.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import this module.

Axis convention: the product's IR stores every leaf in a canonical axis order
(the "output" axis of the generator first, then its inputs); the network
wiring, not the array, knows which axes are dom/cod, so daggers need no
second array (all built-in amplitudes are real, diagram.py:588-590).
"""
from __future__ import annotations

import itertools
from typing import Iterable, Tuple

import numpy as np

# kinds -- must match optyx_b200/core/network.py and include/optyx_b200.h
K_ONEHOT, K_ONES, K_VEC, K_DENSE, K_SUM, K_ADD, K_MUL, K_DIV, K_MOD2, \
    K_DUALRAIL, K_PTD, K_EMBED = range(1, 13)


def occupation_numbers(n_photons: int, m_modes: int):
    """optyx/utils/utils.py:90-111."""
    if not n_photons:
        return [m_modes * (0,)]
    if not m_modes:
        raise ValueError(f"Can't put {n_photons} photons in zero modes!")
    if m_modes == 1:
        return [(n_photons,)]
    return [
        (head,) + tail
        for head in range(n_photons, -1, -1)
        for tail in occupation_numbers(n_photons - head, m_modes - 1)
    ]


def multinomial(lst) -> int:
    """optyx/utils/utils.py:114-124 (exact integer arithmetic)."""
    lst = list(lst)
    res, i = 1, sum(lst)
    i0 = lst.index(max(lst))
    for a in lst[:i0] + lst[i0 + 1:]:
        for j in range(1, a + 1):
            res *= i
            res //= j
            i -= 1
    return res


def _scatter(input_dims, output_dims, spec) -> np.ndarray:
    """``Box.truncation`` scatter loop, diagram.py:559-581: result shape is
    ``(*output_dims, *input_dims)`` and ``A[out + inp] = amp``."""
    arr = np.zeros((*output_dims, *input_dims), dtype=complex)
    ranges = [range(int(d)) for d in input_dims]
    for inp in itertools.product(*ranges):
        for out, amp in spec(tuple(inp), tuple(output_dims)):
            arr[tuple(out) + tuple(inp)] = amp
    return arr


# ---- one generator per reference box ---------------------------------------
def _spec_w(n_legs):
    # W.truncation_specification, zw.py:231-244
    def spec(inp, max_out):
        for config in occupation_numbers(inp[0], n_legs):
            if all(config[i] < max_out[i] for i in range(n_legs)):
                yield tuple(config), multinomial(config) ** 0.5
    return spec


def _spec_add(inp, max_out):
    # Add.truncation_specification, zw.py:676-683
    n = sum(inp)
    if n < int(max_out[0]):
        yield (n,), 1.0


def _spec_mul(inp, max_out):
    # Multiply.truncation_specification, zw.py:720-727
    n = int(np.prod(inp))
    if n < int(max_out[0]):
        yield (n,), 1.0


def _spec_div(inp, max_out):
    # Divide.truncation_specification, zw.py:763-772
    a, b = inp
    if b != 0:
        q, r = divmod(a, b)
        if r == 0 and q < int(max_out[0]):
            yield (int(q),), 1.0


def _spec_mod2(inp, max_out):
    # Mod2.truncation_specification, zw.py:818-825
    yield (inp[0] % 2,), 1.0


def _spec_dualrail(inp, max_out):
    # DualRail.truncation_specification, diagram.py:904-913
    yield ((1, 0) if inp[0] == 0 else (0, 1)), 1.0


def _spec_ptd(inp, max_out):
    # PhotonThresholdDetector.truncation_specification, diagram.py:975-983
    yield ((0,) if inp[0] == 0 else (1,)), 1.0


def build_leaf(kind: int, shape: Tuple[int, ...], ipar: int = 0,
               params=None) -> np.ndarray:
    """Dense complex128 array of one leaf descriptor, canonical axis order."""
    shape = tuple(int(d) for d in shape)
    if kind == K_ONEHOT:
        # Create / Select one-hot (zw.py:450-460, 567-573): zero when the
        # occupation does not fit the truncated wire
        arr = np.zeros(shape, dtype=complex)
        if ipar < shape[0]:
            arr[ipar] = 1.0
        return arr
    if kind == K_ONES:
        return np.ones(shape, dtype=complex)
    if kind in (K_VEC, K_DENSE):
        return np.asarray(params, dtype=complex).reshape(shape).copy()
    if kind == K_SUM:
        # canonical [n, n_1..n_k]: the generator's input is n, outputs n_i;
        # the reference array is A[out + inp] = [n_1..n_k, n]
        arr = _scatter(shape[:1], shape[1:], _spec_w(len(shape) - 1))
        return np.moveaxis(arr, -1, 0)
    if kind == K_ADD:
        # canonical [s, n_1..n_k] == reference A[out + inp] directly
        return _scatter(shape[1:], shape[:1], _spec_add)
    if kind == K_MUL:
        return _scatter(shape[1:], shape[:1], _spec_mul)
    if kind == K_DIV:
        return _scatter(shape[1:], shape[:1], _spec_div)
    if kind == K_MOD2:
        return _scatter(shape[1:], shape[:1], _spec_mod2)
    if kind == K_PTD:
        return _scatter(shape[1:], shape[:1], _spec_ptd)
    if kind == K_DUALRAIL:
        # canonical [b, m0, m1]; reference A[out + inp] = [m0, m1, b]
        arr = _scatter(shape[:1], shape[1:], _spec_dualrail)
        return np.moveaxis(arr, -1, 0)
    if kind == K_EMBED:
        # EmbeddingTensor, diagram.py:1010-1024 (transposed rectangular eye)
        input_dim, output_dim = shape
        emb = np.zeros((output_dim, input_dim), dtype=complex)
        if input_dim < output_dim:
            emb[:input_dim, :input_dim] = np.eye(input_dim)
        else:
            emb[:output_dim, :output_dim] = np.eye(output_dim)
        return emb.T.copy()
    raise ValueError(f"unknown leaf kind {kind}")


def build_all(leaves: Iterable) -> list:
    """Arrays for a list of leaf descriptors (anything with ``kind``,
    ``shape``, ``ipar``, ``params`` attributes)."""
    return [build_leaf(l.kind, l.shape, l.ipar, l.params) for l in leaves]
