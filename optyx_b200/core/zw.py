"""ZW generators: host-side mirror of optyx/core/zw.py.

Each class keeps the reference's constructor, ``dagger``/``conjugate`` rules,
``determine_output_dimensions`` and light-cone flags; ``emit`` replaces
``truncation``/``truncation_specification`` and only records a leaf
*descriptor* -- the dense truncated-Fock array is evaluated on the device.
"""
from __future__ import annotations

from typing import List

import numpy as np

from . import diagram
from . import network as nw
from .diagram import Diagram, Mode, PhotonNumberPreservation, Ty


class ZWBox(diagram.Box):
    """zw.py:165-175"""

    def __init__(self, name, dom, cod, **params):
        if isinstance(dom, int):
            dom = Mode(dom)
        if isinstance(cod, int):
            cod = Mode(cod)
        super().__init__(name=name, dom=dom, cod=cod, **params)
        self.photon_preservation_behaviour = PhotonNumberPreservation.LO


class IndexableAmplitudes:
    """zw.py:178-209"""

    def __init__(self, func, conj=False, name="Z(func)"):
        self.func, self.conj, self.name = func, conj, name

    def __getitem__(self, i):
        if not self.conj:
            return self.func(i)
        return np.conj(self.func(i))

    def conjugate(self):
        return IndexableAmplitudes(self.func, conj=not self.conj, name=self.name)


class W(ZWBox):
    """W node, zw.py:212-272: ``W[n; n_1..n_k] = sqrt(n!/prod n_i!)``."""

    def __init__(self, n_legs: int, is_dagger: bool = False):
        dom = Mode(n_legs) if is_dagger else Mode(1)
        cod = Mode(1) if is_dagger else Mode(n_legs)
        super().__init__("W", dom, cod)
        self.n_legs = n_legs
        self.is_dagger = is_dagger

    def conjugate(self):
        return self

    def dagger(self):
        return W(self.n_legs, not self.is_dagger)

    def determine_output_dimensions(self, input_dims):
        if self.is_dagger:
            return [int(np.sum(np.array(input_dims) - 1) + 1)]
        return [input_dims[0] for _ in range(len(self.cod))]

    def emit(self, b, ins, dims_in, dims_out):
        outs = [b.new_index(d) for d in dims_out]
        # canonical leaf axes [n, n_1..n_k]
        inds = [outs[0]] + list(ins) if self.is_dagger else [ins[0]] + outs
        b.add(nw.K_SUM, inds, name="W")
        return outs

    def __repr__(self):
        return f"W({self.n_legs}{', dagger=True' if self.is_dagger else ''})"


class ZBox(diagram.Spider, ZWBox):
    """Z spider with per-occupation amplitudes, zw.py:275-395."""

    def __init__(self, legs_in: int = 1, legs_out: int = 1, amplitudes=lambda i: i):
        if callable(amplitudes):
            self.amplitudes = IndexableAmplitudes(amplitudes)
        else:
            self.amplitudes = amplitudes
        diagram.Box.__init__(self, "Z", Mode(legs_in), Mode(legs_out))
        self.typ = diagram.mode
        self.legs_in, self.legs_out = legs_in, legs_out
        self.photon_preservation_behaviour = PhotonNumberPreservation.NON_LO

    def conjugate(self):
        # zw.py:299-300 -- calls .conjugate() on whatever the amplitudes are
        return ZBox(self.legs_in, self.legs_out, self.amplitudes.conjugate())

    def dagger(self):
        # zw.py:394-395
        amps = self.amplitudes.conjugate() if isinstance(self.amplitudes, IndexableAmplitudes) \
            else np.conj(self.amplitudes)
        return ZBox(self.legs_out, self.legs_in, amps)

    def determine_output_dimensions(self, input_dims):
        if self.legs_in == 0:
            return [2 for _ in range(len(self.cod))]
        return [min(input_dims) for _ in range(len(self.cod))]

    def _diag(self, spider_dim: int):
        # zw.py:320-329
        if not isinstance(self.amplitudes, IndexableAmplitudes) and \
                spider_dim > len(self.amplitudes):
            return list(self.amplitudes) + [0] * (spider_dim - len(self.amplitudes))
        return [self.amplitudes[i] for i in range(spider_dim)]

    def emit(self, b, ins, dims_in, dims_out):
        if self.legs_out == 0 and self.legs_in == 0:      # zw.py:309-316
            amp = self.amplitudes[0] if isinstance(self.amplitudes, IndexableAmplitudes) \
                else self.amplitudes
            b.mul_scalar(complex(np.array([amp], dtype=complex).reshape(-1)[0]))
            return []
        spider_dim = int(min(dims_in)) if self.legs_in > 0 else 2
        return diagram.emit_spider(b, ins, dims_in, self.legs_out, spider_dim,
                                   amplitudes=self._diag(spider_dim))

    def inflate(self, d):
        return diagram.Box.inflate(self, d)


class Create(ZWBox):
    """zw.py:398-504"""

    def __init__(self, *photons: int, internal_states=None):
        self.photons = photons or (1,)
        if internal_states is not None:
            if not isinstance(internal_states, tuple):
                internal_states = (internal_states,)
            assert all(p in (0, 1) for p in photons), \
                "Only 0 or 1 photons per mode are allowed for internal states"
            assert sum(photons) == len(internal_states), \
                "The number of internal states must match the total number of photons"
            assert len(set(len(i) for i in internal_states)) == 1, \
                "All internal states must be of the same length"
        self.internal_states = internal_states
        name = "Create(1)" if self.photons == (1,) else f"Create({photons})"
        super().__init__(name, 0, len(self.photons))
        # CUSTOM is assigned *before* super().__init__ in the reference
        # (zw.py:433-436), so ZWBox.__init__ leaves the box LO.

    def conjugate(self):
        return self

    def dagger(self):
        return Select(*self.photons, internal_states=self.internal_states)

    def determine_output_dimensions(self, input_dims=None):
        return [p + 1 if p != 0 else 2 for p in self.photons]

    def emit(self, b, ins, dims_in, dims_out):
        outs = [b.new_index(d) for d in dims_out]
        for o, p in zip(outs, self.photons):
            b.add(nw.K_ONEHOT, [o], ipar=int(p), name="Create")
        return outs

    def inflate(self, d):
        # zw.py:476-504
        if any(p == 1 for p in self.photons):
            assert self.internal_states is not None, \
                "Internal states in Create/Select must be provided if there is " \
                "at least one photon being created"
            assert all(len(s) == d for s in self.internal_states), \
                "All internal states must be of length d"
        dgrm = Diagram.id(Mode(0))
        photon_index = 0
        for n_photons in self.photons:
            endo_layer = Diagram.id(Mode(0))
            if n_photons == 1:
                for j in self.internal_states[photon_index]:
                    endo_layer = endo_layer @ Endo(j)
                photon_index += 1
            else:
                endo_layer = endo_layer @ Diagram.id(Mode(d))
            dgrm = dgrm @ (Create(n_photons) >> W(d) >> endo_layer)
        return dgrm


class Select(ZWBox):
    """zw.py:507-580"""

    def __init__(self, *photons: int, internal_states=None):
        if internal_states is not None:
            if not isinstance(internal_states, tuple):
                internal_states = (internal_states,)
            assert all(p in (0, 1) for p in photons)
            assert sum(photons) == len(internal_states)
            assert len(set(len(i) for i in internal_states)) == 1
        self.internal_states = internal_states
        self.photons = photons or (1,)
        name = "Select(1)" if self.photons == (1,) else f"Select({photons})"
        super().__init__(name, len(self.photons), 0)

    def conjugate(self):
        return self

    def dagger(self):
        return Create(*self.photons, internal_states=self.internal_states)

    def determine_output_dimensions(self, _=None):
        return []

    def emit(self, b, ins, dims_in, dims_out):
        for i, p in zip(ins, self.photons):
            b.add(nw.K_ONEHOT, [i], ipar=int(p), name="Select")
        return []

    def inflate(self, d):
        return Create(*self.photons, internal_states=self.internal_states).inflate(d).dagger()


class Endo(ZWBox):
    """zw.py:583-649: ``diag(s ** k)``; the powers are evaluated on the host by
    the same Python expression as the reference (zw.py:644)."""

    def __init__(self, scalar: complex):
        try:
            scalar = complex(scalar)
        except TypeError:
            pass
        self.scalar = scalar
        super().__init__(f"Endo({scalar})", 1, 1, data=scalar)

    def conjugate(self):
        return Endo(self.scalar.conjugate())

    def dagger(self):
        return Endo(self.scalar.conjugate())

    def determine_output_dimensions(self, input_dims):
        return input_dims

    def emit(self, b, ins, dims_in, dims_out):
        d = int(dims_in[0])
        b.add(nw.K_VEC, [ins[0]], params=np.array([self.scalar ** x for x in range(d)],
                                                  dtype=complex), name="Endo")
        return [ins[0]]

    def inflate(self, d):
        return diagram.Box.inflate(self, d)


class _Arith(ZWBox):
    """Shared plumbing of Add / Multiply / Divide / Mod2: canonical leaf axes
    are [result, operands...]; the dagger only swaps which side is dom."""
    kind = 0

    def conjugate(self):
        return self

    def emit(self, b, ins, dims_in, dims_out):
        outs = [b.new_index(d) for d in dims_out]
        inds = [ins[0]] + outs if self.is_dagger else [outs[0]] + list(ins)
        b.add(self.kind, inds, name=self.name)
        return outs


class Add(_Arith):
    """zw.py:652-694"""
    kind = nw.K_ADD

    def __init__(self, n: int, is_dagger: bool = False):
        dom = Mode(1) if is_dagger else Mode(n)
        cod = Mode(n) if is_dagger else Mode(1)
        super().__init__("Add", dom, cod)
        self.n, self.is_dagger = n, is_dagger

    def determine_output_dimensions(self, input_dims):
        if self.is_dagger:
            return [int(input_dims[0])] * self.n
        return [int(sum(input_dims))]

    def dagger(self):
        return Add(self.n, not self.is_dagger)


class Multiply(_Arith):
    """zw.py:697-738"""
    kind = nw.K_MUL

    def __init__(self, is_dagger: bool = False):
        dom = Mode(1) if is_dagger else Mode(2)
        cod = Mode(2) if is_dagger else Mode(1)
        super().__init__("Multiply", dom, cod)
        self.is_dagger = is_dagger
        self.photon_preservation_behaviour = PhotonNumberPreservation.NON_LO

    def determine_output_dimensions(self, input_dims):
        if self.is_dagger:
            return [int(input_dims[0]), int(input_dims[0])]
        return [int(np.prod(input_dims))]

    def dagger(self):
        return Multiply(not self.is_dagger)


class Divide(_Arith):
    """zw.py:741-783"""
    kind = nw.K_DIV

    def __init__(self, is_dagger: bool = False):
        dom = Mode(1) if is_dagger else Mode(2)
        cod = Mode(2) if is_dagger else Mode(1)
        super().__init__("Divide", dom, cod)
        self.is_dagger = is_dagger
        self.photon_preservation_behaviour = PhotonNumberPreservation.NON_LO

    def determine_output_dimensions(self, input_dims):
        if self.is_dagger:
            return [diagram.MAX_DIM, diagram.MAX_DIM]
        return [int(input_dims[0])]

    def dagger(self):
        return Divide(not self.is_dagger)


class Mod2(_Arith):
    """zw.py:786-836"""
    kind = nw.K_MOD2

    def __init__(self, is_dagger: bool = False):
        dom = diagram.Bit(1) if is_dagger else Mode(1)
        cod = Mode(1) if is_dagger else diagram.Bit(1)
        super().__init__("Mod2", dom, cod)
        self.is_dagger = is_dagger
        self.photon_preservation_behaviour = PhotonNumberPreservation.CUSTOM

    def photon_number_transform(self, dims_in, dims_out):
        from .zx import Z
        if self.is_dagger:
            return [Z(1, 0), Create(dims_out[0] - 1)]
        return [Select(0), Z(0, 1)]

    def determine_output_dimensions(self, input_dims):
        return [diagram.MAX_DIM] if self.is_dagger else [2]

    def dagger(self):
        return Mod2(not self.is_dagger)


SWAP = diagram.Swap(diagram.mode, diagram.mode)


def Split(n):
    return W(n)


def Merge(n):
    return W(n).dagger()


def Id(n):
    return Diagram.id(n) if isinstance(n, Ty) else Diagram.id(Mode(n))


LO_ELEMENTS = (Create, W, Select, Add, Endo, diagram.Swap, Add)
