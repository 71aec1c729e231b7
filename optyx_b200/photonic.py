"""Photonic channel generators: host-side mirror of optyx/photonic.py (the
subset that feeds ``eval``: every class only assembles ZW kraus diagrams; all
arithmetic is the leaf kernels').  Symbolic (sympy) parameters, ``grad`` and
``lambdify`` are out of scope: parameters are plain numbers.
"""
from __future__ import annotations

import numpy as np

from .core import channel, diagram, zw
from .core.channel import (Channel, Diagram, Discard as DiscardChannel,
                           Encode as EncodeChannel, Measure, Ty, bit, mode, qmode)


def _build_w_layer(counts, dagger=False):
    """utils.py:20-30"""
    layer = zw.Id(0)
    for count in counts:
        if count > 1:
            w = zw.W(int(count))
            layer = layer @ (w.dagger() if dagger else w)
        elif count == 1:
            layer = layer @ zw.Id(1)
    return layer


def matrix_to_zw(U) -> diagram.Diagram:
    """utils.py:33-87: ``U`` (dom x cod) -> W layer >> Endo(U_rc) per non-zero,
    row major >> regroup by column >> W-dagger layer."""
    U = np.asarray(U)
    n = U.shape[0]
    # the reference counts non-zeros as |sign(U)| (utils.py:45, 69, 84); exact here
    nonzero = (U != 0).astype(int)
    n_cols_nonzero = nonzero.sum(axis=1)
    d = zw.Id(0) @ _build_w_layer(n_cols_nonzero, dagger=False)
    endo_layer = zw.Id(0)
    rows, cols = np.nonzero(U)
    for r, c in zip(rows, cols):
        endo_layer = endo_layer @ zw.Endo(U[r, c])
    d = d >> endo_layer
    row_idx, col_idx = np.nonzero(U)
    order = np.lexsort((row_idx, col_idx))
    swap_list = [int(n * r + c) for r, c in zip(row_idx[order], col_idx[order])]
    flat = nonzero.flatten()
    adjusted = [int(idx - np.abs(np.array(flat)[:idx] - 1).sum()) for idx in swap_list]
    if adjusted:
        d = d.permute(*adjusted)
    n_rows_nonzero = nonzero.sum(axis=0)
    return d >> _build_w_layer(n_rows_nonzero, dagger=True)


class Scalar(Channel):
    """photonic.py:272-306"""

    def __init__(self, value):
        self.scalar = complex(value)
        super().__init__(f"{value}", diagram.Scalar(value))
        self.data = value


class Encode(EncodeChannel):
    def __init__(self, n):
        super().__init__(mode ** n)


class Discard(DiscardChannel):
    def __init__(self, n):
        super().__init__(qmode ** n)


class PhotonThresholdMeasurement(Channel):
    """photonic.py:326-338"""

    def __init__(self, n=1):
        super().__init__("PhotonThresholdMeasurement",
                         diagram.PhotonThresholdDetector() ** n, cod=bit ** n)


class NumberResolvingMeasurement(Measure):
    """photonic.py:341-347"""

    def __init__(self, n):
        super().__init__(qmode ** n)


class Create(Channel):
    """photonic.py:350-360"""

    def __init__(self, *photons: int, internal_states=None):
        self.photons = photons
        self.internal_states = internal_states
        super().__init__(f"Create({photons})",
                         zw.Create(*photons, internal_states=internal_states))


class AbstractGate(Channel):
    """photonic.py:363-401"""

    def __init__(self, dom: int, cod: int, name: str, data=None):
        self.data = data
        super().__init__(name, matrix_to_zw(self.array.reshape(dom, cod)))

    @property
    def array(self):
        if not hasattr(self, "_gate_array"):
            self._gate_array = np.asarray(self._compute_array())
        return self._gate_array

    def _compute_array(self):
        raise NotImplementedError


class Gate(AbstractGate):
    """photonic.py:404-455"""

    def __init__(self, matrix, dom: int, cod: int, name: str, data=None):
        self._matrix = np.asanyarray(matrix)
        super().__init__(dom, cod, name, data=data)

    def _compute_array(self):
        return self._matrix

    def dagger(self):
        return Gate(np.conjugate(self.array.T), len(self.cod), len(self.dom), self.name)

    def conjugate(self):
        return Gate(np.conjugate(self.array), len(self.dom), len(self.cod), self.name)


class Phase(AbstractGate):
    """photonic.py:458-507"""

    def __init__(self, angle: float):
        self.angle = angle
        super().__init__(1, 1, f"Phase({angle})", data=angle)

    def _compute_array(self):
        return [np.exp(2 * np.pi * 1j * self.angle)]

    def dagger(self):
        return Phase(-self.angle)

    def conjugate(self):
        return Phase(-self.angle)


class NumOp(Channel):
    """photonic.py:510-522"""

    def __init__(self):
        super().__init__("NumOp", (zw.Split(2) >> zw.Id(1) @ (zw.Select() >> zw.Create())
                                   >> zw.Merge(2)))


class BBS(AbstractGate):
    """photonic.py:525-605"""

    def __init__(self, bias, is_conj=False):
        self.bias, self.is_conj = bias, is_conj
        super().__init__(2, 2, f"BBS({bias})", data=bias)

    def _compute_array(self):
        sin = np.sin((0.25 + self.bias) * np.pi)
        cos = np.cos((0.25 + self.bias) * np.pi)
        if self.is_conj:
            array = [-1j * cos, sin, sin, -1j * cos]
        else:
            array = [1j * cos, sin, sin, 1j * cos]
        return np.array(array).reshape(2, 2)

    def dagger(self):
        return BBS(0.5 - self.bias)

    def conjugate(self):
        return BBS(self.bias, not self.is_conj)


class TBS(AbstractGate):
    """photonic.py:608-703"""

    def __init__(self, theta, is_gate_dagger=False, is_conj=False):
        self.theta, self.is_gate_dagger, self.is_conj = theta, is_gate_dagger, is_conj
        super().__init__(2, 2, f"TBS({theta})", data=theta)

    @property
    def global_phase(self):
        return (-1j * np.exp(-1j * self.theta * np.pi) if self.is_gate_dagger or self.is_conj
                else 1j * np.exp(1j * self.theta * np.pi))

    def _compute_array(self):
        sin, cos = np.sin(self.theta * np.pi), np.cos(self.theta * np.pi)
        array = np.array([sin, cos, cos, -sin]).reshape(2, 2)
        if self.is_gate_dagger:
            array = np.conjugate(array.T)
        if self.is_conj:
            array = np.conjugate(array)
        return array * self.global_phase

    def conjugate(self):
        return TBS(self.theta, self.is_gate_dagger, not self.is_conj)

    def dagger(self):
        return TBS(self.theta, is_gate_dagger=not self.is_gate_dagger, is_conj=self.is_conj)


class MZI(AbstractGate):
    """photonic.py:706-811"""

    def __init__(self, theta, phi, is_gate_dagger=False, is_conj=False):
        self.theta, self.phi = theta, phi
        self.is_gate_dagger, self.is_conj = is_gate_dagger, is_conj
        super().__init__(2, 2, f"MZI({theta}, {phi})", data=(theta, phi))

    @property
    def global_phase(self):
        return (-1j * np.exp(-1j * self.theta * np.pi) if self.is_gate_dagger or self.is_conj
                else 1j * np.exp(1j * self.theta * np.pi))

    def _compute_array(self):
        cos, sin = np.cos(np.pi * self.theta), np.sin(np.pi * self.theta)
        exp = np.exp(1j * 2 * np.pi * self.phi)
        array = np.array([exp * sin, cos, exp * cos, -sin]).reshape(2, 2)
        if self.is_gate_dagger:
            array = np.conjugate(array.T)
        if self.is_conj:
            array = np.conjugate(array)
        return array * self.global_phase

    def dagger(self):
        return MZI(self.theta, self.phi, is_gate_dagger=not self.is_gate_dagger,
                   is_conj=self.is_conj)

    def conjugate(self):
        return MZI(self.theta, self.phi, self.is_gate_dagger, not self.is_conj)


def ansatz(width, depth, params=None):
    """photonic.py:814-851.  The reference returns sympy symbols ``a_i_j``,
    ``b_i_j`` to substitute; here ``params`` is a dict ``{(i, j): (a, b)}`` or
    a callable ``(i, j) -> (a, b)`` of numbers."""
    def p(i, j):
        if callable(params):
            return params(i, j)
        return params[(i, j)]
    d = Diagram.id(qmode ** width)
    for i in range(depth):
        n_mzi = (width - 1) // 2 if i % 2 else width // 2
        left = qmode ** (i % 2)
        right = qmode ** (width - (i % 2) - 2 * n_mzi)
        wire = Diagram.id(qmode ** 0)
        for j in range(n_mzi):
            wire = wire @ MZI(*p(i, j))
        d = d >> (left @ wire @ right)
    return d


def ansatz_symbol_names(width, depth):
    """Sorted symbol names of the reference's ansatz (``a_i_j``, ``b_i_j``), the
    order in which seeded values are assigned in SURVEY.md 8(d) cfg1."""
    names = []
    for i in range(depth):
        n_mzi = (width - 1) // 2 if i % 2 else width // 2
        for j in range(n_mzi):
            names += [f"a_{i}_{j}", f"b_{i}_{j}"]
    return sorted(names)


def seeded_ansatz(width, depth, seed=0):
    """``ansatz`` with every parameter ~U(0,1) from ``default_rng(seed)`` in
    sorted-symbol order."""
    rng = np.random.default_rng(seed)
    vals = {name: float(rng.uniform(0, 1)) for name in ansatz_symbol_names(width, depth)}
    return ansatz(width, depth, lambda i, j: (vals[f"a_{i}_{j}"], vals[f"b_{i}_{j}"]))


class HadamardBS(Gate):
    """photonic.py:854-864"""

    def __init__(self):
        super().__init__(np.sqrt(1 / 2) * np.array([[1, 1], [1, -1]]), 2, 2, "HadamardBS")


class DualRail(Channel):
    """photonic.py:867-876"""

    def __init__(self, n_qubits, internal_states=None):
        if internal_states is None:
            d = diagram.DualRail()
            for _ in range(n_qubits - 1):
                d = d @ diagram.DualRail()
        else:
            d = diagram.DualRail(internal_state=internal_states[0])
            for state in internal_states[1:]:
                d = d @ diagram.DualRail(internal_state=state)
        super().__init__(f"DualRail({n_qubits})", d)


class Swap(channel.Swap):
    def __init__(self, left, right):
        super().__init__(qmode ** left, qmode ** right)


class PhotonLoss(Channel):
    """photonic.py:1007-1052: ``W(2) >> Endo(sqrt p) @ Endo(sqrt(1-p))``, the
    second leg is the environment."""

    def __init__(self, p_survive):
        self.p_survive = p_survive
        kraus = zw.W(2) >> zw.Endo(p_survive ** 0.5) @ zw.Endo((1 - p_survive) ** 0.5)
        super().__init__(f"Loss({p_survive})", kraus, qmode, qmode @ Ty(), env=diagram.mode)

    def inflate(self, d):
        # the reference's Channel.inflate (channel.py:670-681) drops ``env`` and
        # so cannot inflate a lossy channel; each of the d internal modes loses
        # photons independently: permute (out_k, env_k) pairs apart
        kraus = self.kraus.inflate(d)
        return Channel(self.name + f"^{d}", kraus, qmode ** d, qmode ** d,
                       env=diagram.Mode(d))


def _swap11():
    return diagram.Swap(diagram.mode, diagram.mode)


def fusion_I_function(x):
    """photonic.py:975-984 (module level: a pure function, its truth table is memoised)"""
    a, b = x[0], x[1]
    s = (a % 2) ^ (b % 2)
    k = int(s * b + (1 - s) * (1 - (a + b) / 2)) % 2
    return [s, k]


def fusion_II_function(x):
    """photonic.py:1114-1124"""
    a, b, d = x[0], x[1], x[3]
    s = (a % 2) ^ (b % 2)
    k = int(s * (b + d) + (1 - s) * (1 - (a + b) / 2)) % 2
    return [s, k]


def FusionTypeI():
    """photonic.py:917-996"""
    from .classical import ClassicalFunction
    M = diagram.Mode
    kraus = (M(1) @ _swap11() @ M(1) >> M(1) @ HadamardBS().get_kraus() @ M(1)
             >> M(2) @ _swap11() >> M(1) @ _swap11() @ M(1))
    fusion = Channel("Fusion I", kraus)

    return (fusion >> qmode ** 2 @ NumberResolvingMeasurement(2)
            >> qmode ** 2 @ ClassicalFunction(fusion_I_function, mode ** 2, bit ** 2))


def FusionTypeII():
    """photonic.py:1055-1137"""
    from .classical import ClassicalFunction
    M = diagram.Mode
    h = HadamardBS().get_kraus
    kraus = (h() @ h() >> M(1) @ _swap11() @ M(1) >> M(1) @ h() @ M(1)
             >> M(2) @ _swap11() >> M(1) @ _swap11() @ M(1) >> h() @ M(2))
    fusion = Channel("Fusion II", kraus)

    return (fusion >> NumberResolvingMeasurement(4)
            >> ClassicalFunction(fusion_II_function, mode ** 4, bit ** 2))


BS = BBS(0)


def Id(n):
    return Diagram.id(n) if isinstance(n, channel.Ty) else Diagram.id(qmode ** n)
