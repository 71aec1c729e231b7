"""Feed-forward boxes: host-side mirror of optyx/core/control.py:65-442.

``ClassicalFunctionBox`` / ``BinaryMatrixBox`` call a user Python function per
basis state (control.py:341-361), which necessarily stays on the host: the
truth table travels to the device as a DENSE leaf.  ``BitControlledBox`` is a
leaf whose two slabs are *contracted sub-networks* (control.py:164-191); the
nested contractions run through the same device executor via
``NetworkBuilder.nested_eval`` (set by the backend).
"""
from __future__ import annotations

import itertools
from typing import Callable, List

import numpy as np

from . import diagram
from . import network as nw
from . import zw
from .diagram import Bit, Diagram, Mode, PhotonNumberPreservation


class ClassicalBox(diagram.Box):
    def conjugate(self):
        return self


def is_diagram_LO(d) -> bool:
    """utils.py:513-524"""
    if diagram.is_identity(d):
        return True
    for box in d.boxes:
        if not isinstance(box, zw.LO_ELEMENTS):
            return False
    return True


def compile_padded(d: Diagram, input_dims: List[int], output_dims: List[int],
                   nested_eval=None) -> nw.Network:
    """``d.to_tensor(input_dims) >> truncation_tensor(cod, output_dims)``
    (control.py:171-178; ``truncation_tensor`` diagram.py:1056-1069)."""
    b = nw.NetworkBuilder(nested_eval)
    ins = [b.new_index(int(x)) for x in input_dims]
    outs, cod_dims = diagram.compile_diagram(d, list(input_dims), builder=b, in_inds=list(ins))
    padded = []
    for o, have, want in zip(outs, cod_dims, output_dims):
        if int(have) == int(want):
            padded.append(o)
        else:
            e = b.new_index(int(want))
            b.add(nw.K_EMBED, [o, e], name="Embedding")
            padded.append(e)
    return b.finish(ins, padded)


class BitControlledBox(ClassicalBox):
    """control.py:65-213"""

    def __init__(self, action_box, default_box=None, is_dagger: bool = False):
        if default_box is None:
            default_box = Diagram.id(action_box.dom)
        assert action_box.dom == default_box.dom and action_box.cod == default_box.cod, \
            "action_box and default_box must have the same domain and codomain"
        assert len(action_box.dom) == len(action_box.cod), \
            "action_box must have the same number of inputs and outputs"
        dom = action_box.cod if is_dagger else Bit(1) @ action_box.dom
        cod = Bit(1) @ action_box.cod if is_dagger else action_box.cod
        name = (action_box.name + "_controlled") if hasattr(action_box, "name") \
            else "controlled_box"
        super().__init__(name, dom, cod)
        self.action_box, self.default_box = action_box, default_box
        self.is_dagger = is_dagger
        self.photon_preservation_behaviour = PhotonNumberPreservation.CUSTOM

    def photon_number_transform(self, dims_in, dims_out):
        # control.py:125-133
        if is_diagram_LO(self.action_box) and is_diagram_LO(self.default_box):
            from .zx import Z
            return (Z(1, 0) if not self.is_dagger else Z(0, 1)) @ self.default_box
        self.photon_preservation_behaviour = PhotonNumberPreservation.NON_LO
        return diagram.Box.photon_number_transform(self, dims_in, dims_out)

    def determine_output_dimensions(self, input_dims):
        # control.py:135-153
        sub_in = list(input_dims) if self.is_dagger else list(input_dims[1:])
        a = self.action_box.cod_dims(sub_in)
        d = self.default_box.cod_dims(sub_in)
        dims = [max(x, y) for x, y in zip(a, d)]
        return [2, *dims] if self.is_dagger else dims

    def emit(self, b, ins, dims_in, dims_out):
        # control.py:155-201
        input_dims, output_dims = (dims_out, dims_in) if self.is_dagger else (dims_in, dims_out)
        action_in = [int(x) for x in input_dims[1:]]
        shape = (2, *action_in, *[int(x) for x in output_dims])
        if b.nested_eval is None:
            raise RuntimeError("BitControlledBox needs a nested evaluator on the builder")
        array = np.zeros(shape, dtype=complex)
        array[0] = np.asarray(b.nested_eval(
            compile_padded(self.default_box, action_in, output_dims, b.nested_eval))).reshape(shape[1:])
        array[1] = np.asarray(b.nested_eval(
            compile_padded(self.action_box, action_in, output_dims, b.nested_eval))).reshape(shape[1:])
        outs = [b.new_index(int(x)) for x in dims_out]
        if self.is_dagger:
            # tensor.Box(...).dagger(): swap the dom/cod axis groups and conjugate
            b.add(nw.K_DENSE, outs + list(ins), params=array.conj(), name=self.name)
        else:
            b.add(nw.K_DENSE, list(ins) + outs, params=array, name=self.name)
        return outs

    def dagger(self):
        return BitControlledBox(self.action_box, self.default_box, not self.is_dagger)

    def conjugate(self):
        return BitControlledBox(self.action_box.conjugate(), self.default_box.conjugate(),
                                self.is_dagger)


class ControlledPhaseShift(ClassicalBox):
    """control.py:216-314.  The array factorises as
    ``prod_k P_k[c, n_k] * delta(in_k, out_k)`` with host-computed phase tables
    ``P_k[c, n] = exp(2 pi i f(c)_k) ** n`` (control.py:276-284)."""

    def __init__(self, function: Callable, n_modes: int = 1, n_control_modes: int = 1,
                 is_dagger: bool = False):
        if n_control_modes != 1:
            raise NotImplementedError("only one control mode is supported")
        dom = Mode(n_modes) if is_dagger else Mode(n_modes + n_control_modes)
        cod = Mode(n_modes + n_control_modes) if is_dagger else Mode(n_modes)
        super().__init__("ControlledPhase", dom, cod)
        self.n_modes, self.function = n_modes, function
        self.is_dagger, self.n_control_modes = is_dagger, n_control_modes

    def determine_output_dimensions(self, input_dims):
        if self.is_dagger:
            return [diagram.MAX_DIM] * self.n_control_modes + list(input_dims)
        return list(input_dims[self.n_control_modes:])

    def emit(self, b, ins, dims_in, dims_out):
        n_c = self.n_control_modes
        if self.is_dagger:
            ctrl = b.new_index(int(dims_out[0]))
            data_in, data_in_dims = list(ins), list(dims_in)
            data_out_dims = list(dims_out[n_c:])
        else:
            ctrl = ins[0]
            data_in, data_in_dims = list(ins[n_c:]), list(dims_in[n_c:])
            data_out_dims = list(dims_out)
        c_dim = b.dim(ctrl)
        phases = [list(self.function(np.array([c]))) for c in range(c_dim)]
        outs = []
        for k, (i, d_in, d_out) in enumerate(zip(data_in, data_in_dims, data_out_dims)):
            # the phase table lives on the ORIGINAL (non-dagger) input side
            d_tab = int(d_out) if self.is_dagger else int(d_in)
            table = np.array([[np.exp(2 * np.pi * 1j * phases[c][k]) ** x
                               for x in range(d_tab)] for c in range(c_dim)], dtype=complex)
            if int(d_in) == int(d_out):
                o = i
            else:
                o = b.new_index(int(d_out))
                b.add(nw.K_EMBED, [i, o], name="Embedding")
            b.add(nw.K_DENSE, [ctrl, o if self.is_dagger else i],
                  params=table.conj() if self.is_dagger else table, name="CtrlPhase")
            outs.append(o)
        return ([ctrl] + outs) if self.is_dagger else outs

    def dagger(self):
        return ControlledPhaseShift(self.function, self.n_modes, self.n_control_modes,
                                    not self.is_dagger)

    def conjugate(self):
        # control.py:311-314: the phases are NOT conjugated (reference quirk)
        return ControlledPhaseShift(self.function, self.n_modes, self.n_control_modes,
                                    self.is_dagger)


_TRUTH_TABLES: dict = {}


class ClassicalFunctionBox(ClassicalBox):
    """control.py:317-376"""

    def __init__(self, function, dom, cod, is_dagger: bool = False):
        assert all(d == cod[0] for d in cod), "cod must be either all Mode(n) or all Bit(n)"
        assert all(d == dom[0] for d in dom), "dom must be either all Mode(n) or all Bit(n)"
        super().__init__("F", dom, cod)
        self.function = function
        self.input_size, self.output_size = len(dom), len(cod)
        self.is_dagger = is_dagger

    def _transition(self, inp, max_output_dims):
        out = self.function(inp)
        if out is None:
            return None
        if isinstance(out, (list, tuple)):
            out = tuple(int(x) for x in out)
        else:
            out = (int(out),)
        if any(x < 0 or x >= int(max_output_dims[i]) for i, x in enumerate(out)):
            return None
        return out

    def _truth_table(self, g_in, g_out) -> np.ndarray:
        """Dense 0/1 array of the function, axes (out..., in...): one call of the user
        function per input basis state (control.py:341-371).  Tables of plain
        module-level functions (no closure state) are memoised per dims: a batch of
        structurally identical diagrams asks for the same table once per diagram."""
        fn = self.function
        cacheable = getattr(fn, "__closure__", 1) is None and \
            getattr(fn, "__qualname__", "<locals>").find("<locals>") < 0
        key = (fn, g_in, g_out)
        if cacheable:
            hit = _TRUTH_TABLES.get(key)
            if hit is not None:
                return hit
        table = np.zeros((*g_out, *g_in), dtype=complex)
        for inp in itertools.product(*[range(x) for x in g_in]):
            out = self._transition(tuple(inp), g_out)
            if out is not None:
                table[tuple(out) + tuple(inp)] = 1.0
        if cacheable and len(_TRUTH_TABLES) < 1024:
            table.setflags(write=False)
            _TRUTH_TABLES[key] = table
        return table

    def determine_output_dimensions(self, input_dims):
        if self.cod == Mode(self.output_size):
            return [diagram.MAX_DIM] * self.output_size
        if self.cod == Bit(self.output_size):
            return [2] * self.output_size
        return [int(max(input_dims))] * self.output_size

    def emit(self, b, ins, dims_in, dims_out):
        # Box.truncation (diagram.py:556-590) with this generator: the function
        # runs over the generator's inputs, which for a dagger box are the
        # actual outputs
        g_in, g_out = (dims_out, dims_in) if self.is_dagger else (dims_in, dims_out)
        table = self._truth_table(tuple(int(x) for x in g_in), tuple(int(x) for x in g_out))
        outs = [b.new_index(int(x)) for x in dims_out]
        # table axes are (generator out..., generator in...)
        inds = (list(ins) + outs) if self.is_dagger else (outs + list(ins))
        b.add(nw.K_DENSE, inds, params=table, name="F")
        return outs

    def dagger(self):
        return ClassicalFunctionBox(self.function, self.cod, self.dom, not self.is_dagger)


class BinaryMatrixBox(ClassicalFunctionBox):
    """control.py:379-442"""

    def __init__(self, matrix, is_dagger: bool = False):
        matrix = np.array(matrix)
        if len(matrix.shape) == 1:
            matrix = matrix.reshape(1, -1)
        cod = Bit(len(matrix[0])) if is_dagger else Bit(len(matrix))
        dom = Bit(len(matrix)) if is_dagger else Bit(len(matrix[0]))
        self.matrix = matrix

        def f(x):
            x = np.array(x, dtype=np.uint8).reshape(-1, 1)
            a = np.array(self.matrix, dtype=np.uint8)
            return list(((a @ x) % 2).reshape(1, -1)[0])
        diagram.Box.__init__(self, "LogicalMatrix", dom, cod)
        self.function = f
        self.input_size, self.output_size = len(dom), len(cod)
        self.is_dagger = is_dagger

    def dagger(self):
        return BinaryMatrixBox(self.matrix, not self.is_dagger)

    def conjugate(self):
        return self
