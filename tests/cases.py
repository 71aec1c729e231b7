"""Shared test circuits: the reference's own known-answer cases (restated from
its tests/docstrings) and the BASELINE.json configs.  ``oracle_eval`` compiles a
channel diagram with the product's host logic and contracts it with the CPU
oracle; the ``-m gpu`` tests run the same diagrams through ``B200Backend``."""
from __future__ import annotations

import random

import numpy as np

from oracle.contract import contract_network
from optyx_b200 import classical, photonic, qubits
from optyx_b200.core import channel, diagram, zw, zx
from optyx_b200.core.channel import Channel, Discard, Measure, bit, mode, qmode, qubit
from optyx_b200.workloads import *  # noqa: F401,F403  (the diagram builders live in the package)
from optyx_b200.workloads import EVAL_CASES  # noqa: F401


def oracle_nested(net):
    return contract_network(net)


def compile_channel(d: channel.Diagram):
    """The PRODUCT's host compile (``AbstractBackend._get_discopy_tensor``,
    backends.py:406-415) with nested sub-networks contracted by the CPU oracle: the
    leaf-descriptor network the CUDA path would execute."""
    if d.is_pure:
        return d.get_kraus().to_network(nested_eval=oracle_nested)
    return d.double().to_network(nested_eval=oracle_nested)


def _pad_to(arr: np.ndarray, shape) -> np.ndarray:
    out = np.zeros(shape, dtype=arr.dtype)
    out[tuple(slice(0, s) for s in arr.shape)] = arr
    return out


def _sum_terms(d, one) -> np.ndarray:
    # formal sum (backends.py:473-476 + diagram.py:737-765): every term on its
    # own, zero-padded to the largest extent of each axis, then added
    arrs = [one(t) for t in d.terms]
    shape = tuple(max(a.shape[k] for a in arrs) for k in range(arrs[0].ndim))
    return sum(_pad_to(a, shape) for a in arrs)


def oracle_eval(d) -> np.ndarray:
    """The checker: ``QuimbBackend.eval(d).tensor.array`` restated on the CPU by
    ``oracle/compile.py`` + ``oracle/contract.py`` -- CP doubling, per-box
    conjugation, light-cone truncation (literal backward walk), dense truncation
    arrays and contraction, none of it shared with ``optyx_b200``."""
    from oracle.compile import eval_dense
    if hasattr(d, "terms"):
        return _sum_terms(d, oracle_eval)
    return eval_dense(d)


def product_compile_oracle_eval(d) -> np.ndarray:
    """The product's host compile contracted by the CPU oracle (isolates the host
    compile from the kernels)."""
    if hasattr(d, "terms"):
        return _sum_terms(d, product_compile_oracle_eval)
    if isinstance(d, channel.Diagram):
        net = compile_channel(d)
    else:
        net = d.to_network(nested_eval=oracle_nested)
    return contract_network(net)


def statevector_clifford_t(n, depth, seed=0):
    """Independent check of ``random_clifford_t``: plain gate-by-gate statevector
    simulation (qubit 0 = most significant axis), returns <0..0|C|0..0>."""
    rnd = random.Random(seed)
    psi = np.zeros((2,) * n, dtype=complex)
    psi[(0,) * n] = 1
    G = {"H": np.array([[1, 1], [1, -1]]) / np.sqrt(2), "S": np.diag([1, 1j]),
         "T": np.diag([1, np.exp(1j * np.pi / 4)])}
    for layer in range(depth):
        for q in range(n):
            g = G[rnd.choice("HST")]
            psi = np.moveaxis(np.tensordot(g, psi, axes=([1], [q])), 0, q)
        start = layer % 2
        for p in range((n - start) // 2):
            c, t = start + 2 * p, start + 2 * p + 1
            idx = [slice(None)] * n
            idx[c] = 1
            sub = psi[tuple(idx)]
            psi[tuple(idx)] = np.flip(sub, axis=t - 1 if t > c else t)
    return psi[(0,) * n]


def statevector_clifford_t_torch(n, depth, seed=0, device="cuda"):
    """The same independent gate-by-gate statevector simulation for the BENCH-size
    circuit (30 qubits: a 2^30-entry complex128 state, 17 GB) with plain torch ops on
    the GPU -- a TEST-ONLY checker: no kernel, planner or compiler of optyx_b200 is
    involved.  Qubit 0 = most significant axis; returns <0..0|C|0..0> as a Python complex."""
    import math
    import torch
    rnd = random.Random(seed)
    psi = torch.zeros(1 << n, dtype=torch.complex128, device=device)
    psi[0] = 1.0
    r = 1.0 / math.sqrt(2.0)
    phase = {"S": 1j, "T": complex(math.cos(math.pi / 4), math.sin(math.pi / 4))}
    for layer in range(depth):
        for q in range(n):
            g = rnd.choice("HST")
            v = psi.view(1 << q, 2, 1 << (n - q - 1))
            if g == "H":
                a = v[:, 0, :].clone()
                b = v[:, 1, :]
                v[:, 0, :] = (a + b) * r
                v[:, 1, :] = (a - b) * r          # b is read before it is overwritten
                del a
            else:
                v[:, 1, :] *= phase[g]
        start = layer % 2
        for p in range((n - start) // 2):
            c = start + 2 * p                     # control c, target c + 1: swap |10> and |11>
            v = psi.view(1 << c, 2, 2, 1 << (n - c - 2))
            tmp = v[:, 1, 0, :].clone()
            v[:, 1, 0, :] = v[:, 1, 1, :]
            v[:, 1, 1, :] = tmp
            del tmp
    return complex(psi[0].item())
