"""Classical boxes as CQ maps: host-side mirror of optyx/classical.py:156-721."""
from __future__ import annotations

import numpy as np

from .core import channel, control, diagram, zw, zx
from .core.channel import CQMap, Channel, Diagram, Discard, bit, mode, qmode, qubit


class _BitControlledSingleBox(Channel):
    """classical.py:163-199"""

    def __init__(self, control_gate, default_gate=None, classical=False):
        if isinstance(control_gate, channel.Diagram):
            assert control_gate.is_pure, "The input gates must be pure quantum channels"
            control_single = control_gate.get_kraus()
        else:
            control_single = control_gate
        if isinstance(default_gate, channel.Diagram):
            assert default_gate.is_pure, "The input gates must be pure quantum channels"
            default_single = default_gate.get_kraus()
        else:
            default_single = default_gate
        if control_single.dom[0] == diagram.bit:
            tp = qubit if not classical else bit
        else:
            tp = qmode if not classical else mode
        kraus = control.BitControlledBox(control_single, default_single)
        super().__init__("BitControlledGate", kraus,
                         bit @ tp ** len(control_single.dom), tp ** len(control_single.cod))


def BitControlledGate(diag, default_box=None, is_dagger=False, classical=False):
    """classical.py:201-239: copy the control bit in front of every layer whose
    box has ``dom == cod``, control that box, finally delete the control."""
    if default_box is not None:
        single = _BitControlledSingleBox(diag, default_box, classical=classical)
        return single.dagger() if is_dagger else single
    boxes = []
    last_cod = diag.dom
    for left, box, right in diag.layers():
        if box.cod == box.dom:
            copy = Z(1, 2)
            layers = [
                copy @ left @ box.dom @ right,
                Diagram.permutation(
                    [0, *range(2, 2 + len(left)), 1,
                     *range(2 + len(left), 2 + len(left) + len(box.dom))],
                    bit ** 2 @ left @ box.dom) @ right,
                bit @ left @ _BitControlledSingleBox(box) @ right,
            ]
            boxes.append(Diagram.then(*layers))
        else:
            boxes.append(bit @ (Diagram.id(left) @ box @ Diagram.id(right)))
        last_cod = left @ box.cod @ right
    boxes.append(Z(1, 0) @ last_cod)
    res = Diagram.then(*boxes)
    return res.dagger() if is_dagger else res


class BitControlledPhaseShift(Channel):
    """classical.py:242-262"""

    def __init__(self, function, n_modes=1, n_control_modes=1):
        super().__init__("BitControlledPhaseShift",
                         control.ControlledPhaseShift(function, n_modes, n_control_modes),
                         qmode ** n_modes @ mode ** n_control_modes, mode ** n_modes)


def DiscardBit(n):
    return Discard(bit ** n)


def DiscardMode(n):
    return Discard(mode ** n)


class ClassicalBox(CQMap):
    pass


class Scalar(ClassicalBox):
    def __init__(self, value):
        super().__init__(f"{value}", diagram.Scalar(value), bit ** 0, bit ** 0)


class AddN(ClassicalBox):
    def __init__(self, n):
        super().__init__(f"AddInt({n})", zw.Add(n), mode ** n, mode)


class SubN(ClassicalBox):
    def __init__(self):
        super().__init__("SubInt",
                         (zw.Add(2).dagger() @ diagram.Id(diagram.Mode(1))
                          >> diagram.Id(diagram.Mode(1)) @ diagram.Spider(2, 0, diagram.Mode(1))),
                         mode ** 2, mode)


class MultiplyN(ClassicalBox):
    def __init__(self):
        super().__init__("MultiplyInt", zw.Multiply(), mode ** 2, mode)


class DivideN(ClassicalBox):
    def __init__(self):
        super().__init__("DivideInt", zw.Divide(), mode ** 2, mode)


class Mod2(ClassicalBox):
    def __init__(self):
        super().__init__("ModInt", zw.Mod2(), mode, bit)


class CopyN(ClassicalBox):
    def __init__(self, n):
        super().__init__(f"CopyInt({n})", diagram.Spider(1, n, diagram.Mode(1)), mode, mode ** n)


class SwapN(ClassicalBox):
    def __init__(self):
        super().__init__("SwapInt", diagram.Swap(diagram.Mode(1), diagram.Mode(1)),
                         mode ** 2, mode ** 2)


class PostselectBit(ClassicalBox):
    """classical.py:402-421"""

    def __init__(self, *bits):
        if not all(b in (0, 1) for b in bits):
            raise ValueError("Bits must be a list of 0s and 1s.")
        kraus = zx.X(1, 0, 0.5 ** bits[0])
        for b in bits[1:]:
            kraus = kraus @ zx.X(1, 0, 0.5 ** b)
        kraus = kraus @ diagram.Scalar(1 / np.sqrt(2 ** len(bits)))
        super().__init__(f"PostselectBit({bits})", kraus, bit ** len(bits), bit ** 0)


class PostselectDigit(ClassicalBox):
    def __init__(self, *digits):
        if not all(isinstance(d, int) for d in digits):
            raise ValueError("Digits must be a list of integers.")
        super().__init__(f"PostselectDigit({digits})", zw.Select(*digits),
                         mode ** len(digits), mode ** 0)


class NotBit(ClassicalBox):
    def __init__(self):
        super().__init__("NotBit", zx.X(1, 1, 0.5), bit, bit)


class XorBit(ClassicalBox):
    def __init__(self, n=2):
        super().__init__(f"XorBit({n})", zx.X(n, 1) @ diagram.Scalar(np.sqrt(n)), bit ** n, bit)


class AndBit(ClassicalBox):
    def __init__(self, n=2):
        super().__init__("AndBit", zx.And(n), bit ** 2, bit)


class CopyBit(ClassicalBox):
    def __init__(self, n=2):
        super().__init__(f"CopyBit({n})", zx.Z(1, n), bit, bit ** n)


class SwapBit(ClassicalBox):
    def __init__(self):
        super().__init__("SwapBit", diagram.Swap(diagram.Bit(1), diagram.Bit(1)),
                         bit ** 2, bit ** 2)


class OrBit(ClassicalBox):
    def __init__(self, n=2):
        super().__init__(f"OrBit({n})", zx.Or(n), bit ** n, bit)


class Z(ClassicalBox):
    def __init__(self, n_legs_in, n_legs_out, phase=0):
        super().__init__(f"Z({phase})", zx.Z(n_legs_in, n_legs_out, phase),
                         bit ** n_legs_in, bit ** n_legs_out)


class X(ClassicalBox):
    def __init__(self, n_legs_in, n_legs_out, phase=0):
        super().__init__(f"X({phase})", zx.X(n_legs_in, n_legs_out, phase),
                         bit ** n_legs_in, bit ** n_legs_out)


class H(ClassicalBox):
    def __init__(self):
        super().__init__("H", zx.H, bit, bit)


class ClassicalFunction(ClassicalBox):
    """classical.py:596-629"""

    def __init__(self, function, dom, cod):
        box = control.ClassicalFunctionBox(function, dom.single(), cod.single())
        super().__init__(box.name, box,
                         channel.Ty._make(tuple(channel._CLASSICAL[o] for o in box.dom.inside)),
                         channel.Ty._make(tuple(channel._CLASSICAL[o] for o in box.cod.inside)))


class BinaryMatrix(ClassicalBox):
    """classical.py:632-660"""

    def __init__(self, matrix):
        box = control.BinaryMatrixBox(matrix)
        super().__init__(box.name, box,
                         channel.Ty._make(tuple(channel._CLASSICAL[o] for o in box.dom.inside)),
                         channel.Ty._make(tuple(channel._CLASSICAL[o] for o in box.cod.inside)))


class Select(Channel):
    def __init__(self, *photons: int):
        self.photons = photons
        super().__init__(f"Select({photons})", zw.Select(*photons))


class Digit(ClassicalBox):
    def __init__(self, *photons: int):
        self.photons = photons
        super().__init__(f"Digit({photons})", zw.Create(*photons), mode ** 0,
                         mode ** len(photons))


def Bit(*bits):
    return PostselectBit(*bits).dagger()


CtrlX = Channel("Controlled-X", zx.X(2, 1) @ diagram.Scalar(2 ** 0.5), dom=bit @ qubit, cod=qubit)
CtrlZ = Channel("Controlled-Z",
                (zx.H @ diagram.bit >> zx.Z(2, 1) @ diagram.Scalar(2 ** 0.5)),
                dom=bit @ qubit, cod=qubit)


def Id(n):
    if isinstance(n, channel.Ty):
        return Diagram.id(n)
    raise TypeError(f"Expected a channel.Ty, got {type(n)}")
