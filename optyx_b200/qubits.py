"""Qubit channel generators: host-side mirror of optyx/qubits.py:575-932
(gate -> spider table, Ket/Bra, noise channels).  Circuit importers
(tket / pyzx / graphix / discopy) are out of scope."""
from __future__ import annotations

import numpy as np

from .core import channel, diagram, zx
from .core.channel import (Channel, Diagram, Discard as DiscardChannel,
                           Encode as EncodeChannel, Measure as MeasureChannel, bit, qubit)


class Measure(MeasureChannel):
    def __init__(self, n):
        super().__init__(qubit ** n)


class Discard(DiscardChannel):
    def __init__(self, n):
        super().__init__(qubit ** n)


class Encode(EncodeChannel):
    def __init__(self, n):
        super().__init__(bit ** n)


class Z(Channel):
    """qubits.py:630-720"""

    def __init__(self, n_legs_in, n_legs_out, phase=0):
        super().__init__(f"Z({phase})", zx.Z(n_legs_in, n_legs_out, phase),
                         qubit ** n_legs_in, qubit ** n_legs_out)
        self.data = self.phase = phase

    def dagger(self):
        return type(self)(len(self.cod), len(self.dom), -self.phase)


class X(Channel):
    """qubits.py:724-800"""

    def __init__(self, n_legs_in, n_legs_out, phase=0):
        super().__init__(f"X({phase})", zx.X(n_legs_in, n_legs_out, phase),
                         qubit ** n_legs_in, qubit ** n_legs_out)
        self.data = self.phase = phase

    def dagger(self):
        return type(self)(len(self.cod), len(self.dom), -self.phase)


class H(Channel):
    """qubits.py:804-828"""

    def __init__(self):
        super().__init__("H", zx.H, qubit, qubit)

    def dagger(self):
        return self


class Scalar(Channel):
    """qubits.py:831-855"""

    def __init__(self, value):
        super().__init__(f"Scalar({value})", zx.scalar(value), qubit ** 0, qubit ** 0)
        self.data = value

    def dagger(self):
        return type(self)(np.conj(self.data))


class BitFlipError(Channel):
    """qubits.py:858-876"""

    def __init__(self, prob):
        x_error = zx.X(1, 2) >> zx.Id(1) @ zx.ZBox(1, 1, np.sqrt((1 - prob) / prob)) \
            @ zx.scalar(np.sqrt(prob * 2))
        super().__init__(name=f"BitFlipError({prob})", kraus=x_error, dom=qubit, cod=qubit,
                         env=diagram.Bit(1))

    def dagger(self):
        return self


class DephasingError(Channel):
    """qubits.py:879-900"""

    def __init__(self, prob):
        z_error = (zx.H >> zx.X(1, 2)
                   >> zx.H @ zx.ZBox(1, 1, np.sqrt((1 - prob) / prob))
                   @ zx.scalar(np.sqrt(prob * 2)))
        super().__init__(name=f"DephasingError({prob})", kraus=z_error, dom=qubit, cod=qubit,
                         env=diagram.Bit(1))

    def dagger(self):
        return self


class Ket(Channel):
    """qubits.py:903-912"""

    def __init__(self, value, cod=qubit):
        spider = zx.X if value in (0, 1) else zx.Z
        phase = 0 if value in (0, "+") else 0.5
        super().__init__(f"|{value}>", spider(0, 1, phase) @ diagram.Scalar(1 / np.sqrt(2)),
                         cod=cod)


class Bra(Channel):
    """qubits.py:915-924"""

    def __init__(self, value, dom=qubit):
        spider = zx.X if value in (0, 1) else zx.Z
        phase = 0 if value in (0, "+") else 0.5
        super().__init__(f"<{value}|", spider(1, 0, phase) @ diagram.Scalar(1 / np.sqrt(2)),
                         dom=dom)


def Id(n):
    return Diagram.id(n) if isinstance(n, channel.Ty) else Diagram.id(qubit ** n)


root2 = Scalar(2 ** 0.5)


_GATES: dict = {}


def gate(name: str) -> Diagram:
    """The standard-gate -> spider table, qubits.py:575-592.  Diagrams are immutable
    values: one instance per gate name (a circuit of thousands of gates then holds a
    dozen distinct boxes, which the compile caches recognise by identity)."""
    hit = _GATES.get(name)
    if hit is None:
        hit = _GATES[name] = _make_gate(name)
    return hit


def _make_gate(name: str) -> Diagram:
    table = {
        "H": lambda: H(),
        "Z": lambda: Z(1, 1, 0.5),
        "X": lambda: X(1, 1, 0.5),
        "Y": lambda: Z(1, 1, 0.5) >> X(1, 1, 0.5) @ Scalar(1j),
        "S": lambda: Z(1, 1, 0.25),
        "T": lambda: Z(1, 1, 0.125),
        "CZ": lambda: Z(1, 2) @ Id(1) >> Id(1) @ H() @ Id(1) >> Id(1) @ Z(2, 1) @ root2,
        "CX": lambda: Z(1, 2) @ Id(1) >> Id(1) @ X(2, 1) @ root2,
    }
    return table[name]()
