"""ZX generators on ``bit`` wires: host-side mirror of optyx/core/zx.py:285-560.

A Z spider is a shared index plus (when the phase is non-trivial) a length-2
amplitude vector; an X spider is a Z spider with a Hadamard on every leg
(zx.py:406-418).  Reference quirk kept on purpose: ``zx.ZBox`` inherits
``Spider.conjugate`` (zx.py:337-338), which *negates* the "phase" -- for a
ZBox that is the amplitude itself (test/test_qubit.py:107-115 pins it).
"""
from __future__ import annotations

from typing import List

import numpy as np

from . import diagram
from . import network as nw
from . import zw
from .diagram import Bit, Diagram, PhotonNumberPreservation, Ty


class ZXBox(diagram.Box):
    """zx.py:285-320"""

    def __init__(self, name, dom, cod, **params):
        if isinstance(dom, int):
            dom = Bit(dom)
        if isinstance(cod, int):
            cod = Bit(cod)
        super().__init__(name=name, dom=dom, cod=cod, **params)
        self.photon_preservation_behaviour = PhotonNumberPreservation.QUBIT

    def conjugate(self):
        raise NotImplementedError

    def determine_output_dimensions(self, input_dims):
        return [2 for _ in range(len(self.cod))]

    def inflate(self, d):
        return self


class Spider(ZXBox):
    """zx.py:323-374"""

    def __init__(self, n_legs_in, n_legs_out, phase=0):
        shown = "" if np.ndim(phase) == 0 and not phase else \
            (f", {phase}" if np.ndim(phase) == 0 else f", <{np.size(phase)} phases>")
        super().__init__(f"{type(self).__name__}({n_legs_in}, {n_legs_out}{shown})",
                         n_legs_in, n_legs_out)
        self.phase = phase
        self.n_legs_in, self.n_legs_out = n_legs_in, n_legs_out

    def conjugate(self):
        return type(self)(self.n_legs_in, self.n_legs_out, -self.phase)

    def dagger(self):
        return type(self)(len(self.cod), len(self.dom), -self.phase)


def _emit_z(b, ins, n_out, amplitudes) -> List[int]:
    """``zw.ZBox(n, m, [1, a]).truncation([2] * n)`` (zx.py:382-397).  The
    vector is skipped when it is exactly [1, 1] (identity diagonal)."""
    # array-valued phases (one compile for a batch of evals): [2, batch]
    amps = np.stack([np.asarray(x, dtype=complex) for x in np.broadcast_arrays(*amplitudes)])
    # a 1 -> 1 spider is a phase GATE: keep its vector even at phase 0 so that a
    # batch of evals over the phase has one network structure (eval_batch)
    trivial = amps.ndim == 1 and amps[0] == 1 and amps[1] == 1 and not (len(ins) == 1 and n_out == 1)
    return diagram.emit_spider(b, ins, [2] * len(ins), n_out, 2,
                               amplitudes=None if trivial else amps)


class ZBox(Spider):
    """zx.py:377-385: amplitudes ``[1, phase]`` ("phase" is an amplitude)."""

    def emit(self, b, ins, dims_in, dims_out):
        if self.n_legs_in == 0 and self.n_legs_out == 0:
            return zw.ZBox(0, 0, [1, self.phase]).emit(b, ins, dims_in, dims_out)
        return _emit_z(b, ins, self.n_legs_out, [1, self.phase])


class Z(Spider):
    """zx.py:388-397"""

    def emit(self, b, ins, dims_in, dims_out):
        amps = [1, np.exp(1j * self.phase * 2 * np.pi)]
        if self.n_legs_in == 0 and self.n_legs_out == 0:
            return zw.ZBox(0, 0, amps).emit(b, ins, dims_in, dims_out)
        return _emit_z(b, ins, self.n_legs_out, amps)


H_ARRAY = np.array([[1, 1], [1, -1]]) / 2 ** 0.5          # zx.py:437


class X(Spider):
    """zx.py:400-418: ``H^n >> Z(n, m, phase) >> H^m``."""

    def emit(self, b, ins, dims_in, dims_out):
        inner = []
        for i in ins:
            h = b.new_index(2)
            b.add(nw.K_DENSE, [i, h], params=H_ARRAY, name="H")
            inner.append(h)
        amps = [1, np.exp(1j * self.phase * 2 * np.pi)]
        if self.n_legs_in == 0 and self.n_legs_out == 0:
            return zw.ZBox(0, 0, amps).emit(b, ins, dims_in, dims_out)
        mids = _emit_z(b, inner, self.n_legs_out, amps)
        outs = []
        for m in mids:
            o = b.new_index(2)
            b.add(nw.K_DENSE, [m, o], params=H_ARRAY, name="H")
            outs.append(o)
        return outs


def scalar(data):
    return diagram.Scalar(data)


def Id(n):
    return Diagram.id(n) if isinstance(n, Ty) else Diagram.id(Bit(n))


class _Hadamard(ZXBox):
    """zx.py:431-439"""

    def __init__(self):
        super().__init__("H", 1, 1)
        self._array = H_ARRAY

    def conjugate(self):
        return self

    def dagger(self):
        return self

    def emit(self, b, ins, dims_in, dims_out):
        o = b.new_index(2)
        b.add(nw.K_DENSE, [ins[0], o], params=H_ARRAY, name="H")
        return [o]


H = _Hadamard()
SWAP = diagram.Swap(diagram.bit, diagram.bit)       # zx.py:441-449 (a relabel here)


class And(ZXBox):
    """zx.py:452-505"""

    def __init__(self, n=2, is_dagger: bool = False):
        super().__init__("And", Bit(n), Bit(1))
        self.n = n
        self.is_dagger = is_dagger

    def _table(self, n_in):
        array = np.zeros((2,) * n_in + (2,), dtype=complex)
        array[..., 0] = 1
        ones = (1,) * n_in
        array[ones + (0,)] = 0
        array[ones + (1,)] = 1
        return array

    def determine_output_dimensions(self, input_dims):
        return [2] * len(input_dims) if self.is_dagger else [2]

    def emit(self, b, ins, dims_in, dims_out):
        outs = [b.new_index(2) for _ in dims_out]
        b.add(nw.K_DENSE, list(ins) + outs, params=self._table(len(ins)), name=self.name)
        return outs

    def dagger(self):
        return And(not self.is_dagger)           # zx.py:501-502 (sic)

    def conjugate(self):
        return And(n=self.n, is_dagger=self.is_dagger)


class Or(And):
    """zx.py:508-560"""

    def __init__(self, n: int = 2, is_dagger: bool = False):
        ZXBox.__init__(self, "Or", Bit(n), Bit(1))
        self.n = n
        self.is_dagger = is_dagger

    def _table(self, n_in):
        array = np.zeros((2,) * n_in + (2,), dtype=complex)
        array[..., 1] = 1
        zeros = (0,) * n_in
        array[zeros + (1,)] = 0
        array[zeros + (0,)] = 1
        return array

    def dagger(self):
        return Or(n=self.n, is_dagger=not self.is_dagger)

    def conjugate(self):
        return Or(n=self.n, is_dagger=self.is_dagger)
