"""ORACLE (test infrastructure, NOT product code) -- permanents.

Independent second algorithm for *pure linear-optical* circuits: the reference
itself cross-checks QuimbBackend against its PermanentBackend to 1e-12
(test/test_backends.py:482-509).  ``npperm`` restates optyx/core/path.py:141-163
(Glynn formula with Gray code); ``lo_amplitude`` restates the amplitude
formula of ``Matrix.eval`` (path.py:345-384):
    <out| U |in> = perm(U[in-rows repeated, out-cols repeated]) / sqrt(prod in! out!)
for a ``dom x cod`` single-particle matrix ``U``.
"""
from __future__ import annotations

from math import factorial

import numpy as np


def npperm(matrix):
    """path.py:141-163"""
    n = matrix.shape[0]
    if n == 0:
        return 1.0 + 0j
    d = np.ones(n)
    j = 0
    s = 1
    f = np.arange(n)
    v = matrix.sum(axis=0)
    p = np.prod(v)
    while j < n - 1:
        v = v - 2 * d[j] * matrix[j]
        d[j] = -d[j]
        s = -s
        prod = np.prod(v)
        p += s * prod
        f[0] = 0
        f[j] = f[j + 1]
        f[j + 1] = j + 1
        j = f[0]
    return p / 2 ** (n - 1)


def lo_amplitude(U: np.ndarray, creations, selections) -> complex:
    """path.py:361-381 with no open wires: rows = created photons, columns =
    selected photons."""
    U = np.asarray(U, dtype=complex)
    if sum(creations) != sum(selections):
        return 0.0 + 0j
    matrix = np.stack([U[:, m] for m, n in enumerate(selections) for _ in range(n)], axis=1) \
        if sum(selections) else np.zeros((U.shape[0], 0), dtype=complex)
    matrix = np.stack([matrix[m] for m, n in enumerate(creations) for _ in range(n)], axis=0) \
        if sum(creations) else np.zeros((0, 0), dtype=complex)
    divisor = np.sqrt(np.prod([factorial(n) for n in tuple(creations) + tuple(selections)]))
    return npperm(matrix) / divisor
