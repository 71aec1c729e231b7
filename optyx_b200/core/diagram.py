"""Single-picture diagrams (ZW / ZX / classical boxes on ``bit`` and ``mode``
wires) and their compilation to a tensor network.

Host-side mirror of optyx/core/diagram.py.  The reference builds on DisCoPy's
monoidal diagrams (not available here); this module carries the minimal
algebra the hot path needs -- types, boxes, ``>>``, ``@``, dagger, per-box
conjugation, permutations -- and restates ``Diagram.to_tensor``
(diagram.py:301-375) together with the light-cone truncation walker
(optyx/utils/utils.py:347-510).  Instead of a ``discopy.tensor.Diagram`` the
result is an ``optyx_b200.core.network.Network`` of leaf *descriptors*: the
dense arrays are built on the device (csrc/leaf_build.cu).
"""
from __future__ import annotations

from enum import Enum
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import network as nw

MAX_DIM = 10            # diagram.py:217
UNCAPPED = 1e20         # utils.py:446-450


class PhotonNumberPreservation(Enum):
    """diagram.py:220-237"""
    LO = "lo"
    NON_LO = "non_lo"
    CUSTOM = "custom"
    QUBIT = "qubit"


class Ty:
    """A tuple of atomic wire names ("bit" / "mode")."""

    def __init__(self, *obs: str):
        flat: List[str] = []
        for o in obs:
            if isinstance(o, Ty):
                flat.extend(o.inside)
            else:
                flat.append(str(o))
        self.inside: Tuple[str, ...] = tuple(flat)

    def tensor(self, *others: "Ty") -> "Ty":
        res = list(self.inside)
        for o in others:
            res.extend(o.inside)
        return type(self)._make(res)

    @classmethod
    def _make(cls, obs) -> "Ty":
        t = cls.__new__(cls)
        t.inside = tuple(obs)
        return t

    def __matmul__(self, other):
        if isinstance(other, Diagram):
            return other.diagram_factory.id(self) @ other
        return type(self)._make(self.inside + other.inside)

    def __pow__(self, n: int) -> "Ty":
        return type(self)._make(self.inside * n)

    def __len__(self):
        return len(self.inside)

    def __iter__(self):
        for o in self.inside:
            yield type(self)._make((o,))

    def __getitem__(self, key):
        if isinstance(key, slice):
            return type(self)._make(self.inside[key])
        return type(self)._make((self.inside[key],))

    def __eq__(self, other):
        return isinstance(other, Ty) and self.inside == other.inside

    def __hash__(self):
        return hash(self.inside)

    def __repr__(self):
        return " @ ".join(self.inside) if self.inside else "Ty()"

    @property
    def name(self):
        return repr(self)


def Mode(n: int = 0) -> Ty:
    """diagram.py:251-257"""
    return Ty._make(("mode",) * n)


def Bit(n: int = 0) -> Ty:
    """diagram.py:260-266"""
    return Ty._make(("bit",) * n)


bit = Bit(1)
mode = Mode(1)


_SWAP_CACHE: dict = {}
_PERM_CACHE: dict = {}


class Diagram:
    """A list of ``(box, offset)`` layers between ``dom`` and ``cod``."""

    diagram_factory = None      # set below: the class `>>`/`@` build
    ty_factory = Ty

    def __init__(self, dom: Ty, cod: Ty, boxes: Sequence["Box"] = (),
                 offsets: Sequence[int] = ()):
        self.dom, self.cod = dom, cod
        self.boxes: List[Box] = list(boxes)
        self.offsets: List[int] = list(offsets)

    @classmethod
    def _swap_box(cls, left: Ty, right: Ty) -> "Diagram":
        return Swap(left, right)

    @classmethod
    def _spider_box(cls, n_in: int, n_out: int, typ: Ty) -> "Diagram":
        return Spider(n_in, n_out, typ)

    # ---- algebra ------------------------------------------------------------
    @classmethod
    def id(cls, ty=None) -> "Diagram":
        ty = cls.ty_factory() if ty is None else ty
        return cls.diagram_factory(ty, ty)

    def then(self, *others: "Diagram") -> "Diagram":
        res = self
        for o in others:
            res = res >> o
        return res

    def __rshift__(self, other: "Diagram") -> "Diagram":
        if isinstance(other, Sum):
            return NotImplemented               # -> Sum.__rrshift__ (distributes)
        if isinstance(other, Ty):
            other = self.id(other)
        if self.cod != other.dom:
            raise ValueError(f"cannot compose {self.cod} with {other.dom}")
        return self.diagram_factory(self.dom, other.cod, self.boxes + other.boxes,
                                    self.offsets + other.offsets)

    def __lshift__(self, other):
        return other >> self

    def __matmul__(self, other) -> "Diagram":
        if isinstance(other, Sum):
            return NotImplemented               # -> Sum.__rmatmul__
        if isinstance(other, Ty):
            other = self.id(other)
        # f @ g = f @ g.dom >> f.cod @ g   (monoidal whiskering order)
        shift = len(self.cod)
        return self.diagram_factory(self.dom @ other.dom, self.cod @ other.cod,
                                    self.boxes + other.boxes,
                                    self.offsets + [o + shift for o in other.offsets])

    def tensor(self=None, *diagrams: "Diagram") -> "Diagram":
        """``Diagram.tensor(a, b, c)`` or ``a.tensor(b, c)``."""
        ds = ([self] if self is not None else []) + list(diagrams)
        if not ds:
            return Diagram.id(Ty())
        res = ds[0]
        for d in ds[1:]:
            res = res @ d
        return res

    def __pow__(self, n: int) -> "Diagram":
        if n == 1:
            return self
        return self @ self ** (n - 1)

    def __len__(self):
        return len(self.boxes)

    def layers(self):
        """Yield ``(left, box, right)`` per layer."""
        cur = self.dom
        for box, off in zip(self.boxes, self.offsets):
            left, right = cur[:off], cur[off + len(box.dom):]
            yield left, box, right
            cur = left @ box.cod @ right

    def dagger(self) -> "Diagram":
        boxes, offsets = [], []
        for box, off in zip(reversed(self.boxes), reversed(self.offsets)):
            d = box.dagger()
            boxes.extend(d.boxes)
            offsets.extend(o + off for o in d.offsets)
        return self.diagram_factory(self.cod, self.dom, boxes, offsets)

    def _apply_functor(self, on_ob, on_box, factory=None) -> "Diagram":
        """``F(self)`` for a monoidal functor given on objects and boxes, built in one
        pass: layer ``left @ box @ right`` becomes ``F(box)`` whiskered at offset
        ``len(F(left))`` (the same diagram as composing ``id @ F(box) @ id`` layer by
        layer, without the quadratic list copies)."""
        factory = factory or self.diagram_factory
        boxes, offsets = [], []
        # a monoidal functor maps a type wire by wire: track, per wire of the current
        # layer, how many wires its image has, so that ``len(F(left))`` is a prefix sum
        # instead of a type built per layer
        mk = type(self.dom)._make
        width_of: dict = {}

        def widths(ty):
            res = []
            for o in ty.inside:
                w = width_of.get(o)
                if w is None:
                    w = width_of[o] = len(on_ob(mk((o,))))
                res.append(w)
            return res
        cur = widths(self.dom)
        for box, off in zip(self.boxes, self.offsets):
            image = on_box(box)
            shift = sum(cur[:off])
            boxes.extend(image.boxes)
            offsets.extend([o + shift for o in image.offsets] if shift else image.offsets)
            cur[off:off + len(box.dom.inside)] = widths(box.cod)
        return factory(on_ob(self.dom), on_ob(self.cod), boxes, offsets)

    def conjugate(self) -> "Diagram":
        """Per-box conjugation functor, diagram.py:277-284."""
        images = [box.conjugate() for box in self.boxes]
        if all(im.dom == box.dom and im.cod == box.cod for im, box in zip(images, self.boxes)):
            it = iter(images)
            return self._apply_functor(lambda ty: ty, lambda box: next(it))
        res = self.id(self.dom)           # a box whose conjugate changes type (diagram.py:506-519)
        for left, box, right in self.layers():
            res = res >> (self.id(left) @ box.conjugate() @ self.id(right))
        return res

    def inflate(self, d: int) -> "Diagram":
        """diagram.py:420-440"""
        assert isinstance(d, int) and d > 0

        def ob(x: Ty) -> Ty:
            return Ty._make(sum(((o,) * d if o == "mode" else (o,) for o in x.inside), ()))
        return self._apply_functor(ob, lambda box: box.inflate(d), factory=Diagram)

    # ---- symmetric / frobenius structure -------------------------------------
    @classmethod
    def swap(cls, left: Ty, right: Ty) -> "Diagram":
        """Block swap ``left @ right -> right @ left`` out of adjacent ``Swap``s.
        Memoised: diagrams are immutable values and CP doubling / inflation ask for
        the same few swaps thousands of times per compile."""
        key = (cls, type(left), left.inside, right.inside)
        hit = _SWAP_CACHE.get(key)
        if hit is not None:
            return hit
        res = cls._swap_uncached(left, right)
        if len(_SWAP_CACHE) < 4096:
            _SWAP_CACHE[key] = res
        return res

    @classmethod
    def _swap_uncached(cls, left: Ty, right: Ty) -> "Diagram":
        mk = type(left)._make
        res = cls.id(left @ right)
        wires = list(left.inside) + list(right.inside)
        n_l = len(left)
        for j in range(len(right)):
            # move wire n_l + j leftwards to position j
            for p in range(n_l + j, j, -1):
                a, b = wires[p - 1], wires[p]
                res = res >> (cls.id(mk(wires[:p - 1])) @ cls._swap_box(mk((a,)), mk((b,)))
                              @ cls.id(mk(wires[p + 1:])))
                wires[p - 1], wires[p] = b, a
        return res

    @classmethod
    def permutation(cls, xs: Sequence[int], dom: Ty) -> "Diagram":
        """``cod[j] = dom[xs[j]]`` (DisCoPy's ``symmetric.Diagram.permutation``,
        used at channel.py:623-625, 635-637 and utils.py:77-78)."""
        xs = list(xs)
        key = (cls, type(dom), tuple(xs), dom.inside)
        hit = _PERM_CACHE.get(key)
        if hit is not None:
            return hit
        res = cls._permutation_uncached(xs, dom)
        if len(_PERM_CACHE) < 4096:
            _PERM_CACHE[key] = res
        return res

    @classmethod
    def _permutation_uncached(cls, xs, dom: Ty) -> "Diagram":
        assert sorted(xs) == list(range(len(dom))), "not a permutation"
        if len(dom) <= 1:
            return cls.id(dom)
        i = xs[0]
        head = cls.swap(dom[:i], dom[i]) @ cls.id(dom[i + 1:])
        rest_dom = dom[:i] @ dom[i + 1:]
        rest = cls.permutation([x - 1 if x > i else x for x in xs[1:]], rest_dom)
        return head >> (cls.id(dom[i]) @ rest)

    def permute(self, *xs: int) -> "Diagram":
        return self >> self.permutation(list(xs), self.cod)

    @classmethod
    def spiders(cls, n_in: int, n_out: int, typ: Ty) -> "Diagram":
        """Spiders on a compound type: one spider per atomic wire, legs
        interleaved (DisCoPy ``frobenius.Diagram.spiders``; channel.py:629-633)."""
        k = len(typ)
        if k == 1:
            return cls._spider_box(n_in, n_out, typ)
        if k == 0:
            return cls.id(typ)
        # dom = typ ** n_in: wire (copy c, ob j) sits at c*k + j; regroup by ob
        dom = typ ** n_in
        perm_in = [c * k + j for j in range(k) for c in range(n_in)]
        top = cls.permutation(perm_in, dom)
        mid = cls.id(typ[:0])
        for j in range(k):
            mid = mid @ cls._spider_box(n_in, n_out, typ[j])
        grouped = type(typ)._make(sum(((typ.inside[j],) * n_out for j in range(k)), ()))
        perm_out = [j * n_out + c for c in range(n_out) for j in range(k)]
        bot = cls.permutation(perm_out, grouped)
        return top >> mid >> bot

    # ---- compilation ----------------------------------------------------------
    def to_network(self, input_dims: Optional[List[int]] = None,
                   nested_eval=None) -> nw.Network:
        net, _ = compile_diagram(self, input_dims, nested_eval=nested_eval)
        return net

    def to_tensor(self, input_dims: Optional[List[int]] = None, nested_eval=None) -> nw.Network:
        """Name kept from the reference (diagram.py:301-375)."""
        return self.to_network(input_dims, nested_eval=nested_eval)

    def cod_dims(self, input_dims: Optional[List[int]] = None) -> List[int]:
        _, dims = compile_diagram(self, input_dims)
        return dims


Diagram.diagram_factory = Diagram


class Sum:
    """Formal sum of diagrams with one ``dom`` / ``cod`` (diagram.py:710-765).
    ``>>`` / ``@`` distribute over the terms, as DisCoPy's ``symmetric.Sum`` does;
    evaluation contracts every term and adds the results after padding each
    output wire to the per-wire maximum dimension (``Sum.to_tensor``, 737-765)."""

    def __init__(self, terms: Sequence[Diagram], dom: Optional[Ty] = None, cod: Optional[Ty] = None):
        self.terms: List[Diagram] = list(terms)
        if self.terms:
            dom = self.terms[0].dom if dom is None else dom
            cod = self.terms[0].cod if cod is None else cod
        if dom is None or cod is None:
            raise ValueError("an empty sum needs explicit dom and cod")
        for t in self.terms:
            if t.dom != dom or t.cod != cod:
                raise ValueError(f"cannot add diagrams of types {t.dom} -> {t.cod} and {dom} -> {cod}")
        self.dom, self.cod = dom, cod

    def _make(self, terms, dom=None, cod=None):
        return type(self)(terms, dom, cod)

    def __iter__(self):
        return iter(self.terms)

    def __len__(self):
        return len(self.terms)

    def __add__(self, other):
        if isinstance(other, int) and other == 0:
            return self
        others = other.terms if isinstance(other, Sum) else [other]
        return self._make(self.terms + list(others), self.dom, self.cod)

    __radd__ = __add__

    def __rshift__(self, other):
        others = other.terms if isinstance(other, Sum) else [other]
        out = [a >> b for a in self.terms for b in others]
        cod = others[0].cod if others else self.cod
        return self._make(out, self.dom, cod)

    def __rrshift__(self, other):           # diagram >> sum
        return self._make([other >> t for t in self.terms], other.dom, self.cod)

    def __matmul__(self, other):
        others = other.terms if isinstance(other, Sum) else [other]
        out = [a @ b for a in self.terms for b in others]
        return self._make(out, self.dom @ others[0].dom, self.cod @ others[0].cod)

    def __rmatmul__(self, other):
        return self._make([other @ t for t in self.terms], other.dom @ self.dom, other.cod @ self.cod)

    def conjugate(self):
        return self._make([t.conjugate() for t in self.terms], self.dom, self.cod)

    def dagger(self):
        return self._make([t.dagger() for t in self.terms], self.cod, self.dom)

    def inflate(self, d: int):
        terms = [t.inflate(d) for t in self.terms]
        return self._make(terms, terms[0].dom, terms[0].cod) if terms else self

    def to_networks(self, input_dims: Optional[List[int]] = None, nested_eval=None) -> List[nw.Network]:
        """One network per term, output wires padded to the per-wire maximum
        with ``EmbeddingTensor`` leaves (diagram.py:739-757)."""
        nets = [t.to_network(input_dims, nested_eval=nested_eval) for t in self.terms]
        return nw.pad_outputs_to_common_shape(nets)

    def to_tensor(self, input_dims: Optional[List[int]] = None, nested_eval=None):
        return self.to_networks(input_dims, nested_eval=nested_eval)


def _diagram_add(self, other):
    if isinstance(other, int) and other == 0:          # sum([...]) starts from 0
        return self
    if isinstance(other, Sum):
        return other._make([self] + other.terms, self.dom, self.cod)
    return self.sum_factory([self, other], self.dom, self.cod)


Diagram.sum_factory = Sum
Diagram.__add__ = _diagram_add
Diagram.__radd__ = _diagram_add


def is_identity(d: Diagram) -> bool:
    """utils.py:527-528"""
    return len(d.boxes) == 0 and len(d.offsets) == 0


Id = Diagram.id


class Box(Diagram):
    """A generator.  Subclasses provide ``determine_output_dimensions`` and
    ``emit`` (the counterpart of ``truncation``, diagram.py:536-590)."""

    def __init__(self, name: str, dom: Ty, cod: Ty, array=None, **params):
        self.name = name
        self._array = None if array is None else np.asarray(array)
        self.data = params.get("data")
        Diagram.__init__(self, dom, cod, [self], [0])
        self.photon_preservation_behaviour = PhotonNumberPreservation.NON_LO   # :451
        self.is_dagger = False

    # -- light-cone bookkeeping (diagram.py:453-485) ----------------------------
    def photon_number_transform(self, dims_in, dims_out) -> list:
        beh = self.photon_preservation_behaviour
        if beh in (PhotonNumberPreservation.LO, PhotonNumberPreservation.QUBIT):
            return [self]
        if beh == PhotonNumberPreservation.NON_LO:
            from .zw import Create, Select
            res = []
            if len(dims_in) > 0:
                res.append(Select(*[int(i) - 1 for i in dims_in]))
            if len(dims_out) > 0:
                res.append(Create(*[int(i) - 1 for i in dims_out]))
            return res
        raise NotImplementedError(
            f"{type(self).__name__} does not implement photon_number_transform method")

    @staticmethod
    def get_perm(n, d):
        return sorted(sorted(list(range(n))), key=lambda i: i % d)     # :487-489

    def inflate(self, d: int) -> Diagram:
        """diagram.py:491-504"""
        def inflated(ty):
            return Ty._make(ty.inside * d)
        return (Diagram.permutation(self.get_perm(len(self.dom) * d, d), inflated(self.dom))
                >> self ** d
                >> Diagram.permutation(self.get_perm(len(self.cod) * d, d),
                                       inflated(self.cod)).dagger())

    def conjugate(self) -> "Box":
        """diagram.py:506-519 (array boxes swap dom/cod without transposing)."""
        if self._array is not None:
            return type(self)(self.name + ".dagger()", dom=self.cod, cod=self.dom,
                              array=self._array.conjugate())
        raise NotImplementedError(f"{type(self).__name__} does not support conjugation")

    def dagger(self) -> "Box":
        """diagram.py:521-534"""
        if self._array is not None:
            return type(self)(self.name + ".dagger()", dom=self.cod, cod=self.dom,
                              array=self._array.T.conjugate())
        raise NotImplementedError(f"{type(self).__name__} does not support dagger")

    def determine_output_dimensions(self, input_dims: List[int]) -> List[int]:
        if self._array is not None:
            return input_dims                                          # :598-599
        raise NotImplementedError(
            f"{type(self).__name__} does not support determine_output_dimensions")

    @property
    def array(self):
        return self._array

    @array.setter
    def array(self, value):
        self._array = None if value is None else np.asarray(value)

    # -- network emission ---------------------------------------------------------
    def emit(self, b: nw.NetworkBuilder, ins: List[int], dims_in: List[int],
             dims_out: List[int]) -> List[int]:
        """Add this box's leaves; return the output wire indices.  Default:
        an array box on dimension-2 wires (diagram.py:542-548)."""
        if self._array is None:
            raise NotImplementedError(f"{type(self).__name__} has no truncation")
        outs = [b.new_index(2) for _ in self.cod]
        b.add(nw.K_DENSE, list(ins) + outs, params=np.asarray(self._array, dtype=complex),
              name=self.name)
        return outs

    def __repr__(self):
        return self.name


class Spider(Box):
    """Phase-free spider on a single wire type, diagram.py:656-707."""

    def __init__(self, n_legs_in: int, n_legs_out: int, typ: Ty = None, **params):
        typ = bit if typ is None else typ
        assert len(typ) == 1
        self.typ = typ
        super().__init__(f"Spider({n_legs_in}, {n_legs_out}, {typ})", typ ** n_legs_in,
                         typ ** n_legs_out, **params)
        self.photon_preservation_behaviour = PhotonNumberPreservation.NON_LO

    def conjugate(self):
        return self

    def dagger(self):
        return Spider(len(self.cod), len(self.dom), self.typ)

    def inflate(self, d):
        return Box.inflate(self, d)

    def determine_output_dimensions(self, input_dims):
        if self.typ == bit:
            return [2] * len(self.cod)
        if len(self.dom) == 0:
            return [2 for _ in range(len(self.cod))]
        return [min(input_dims) for _ in range(len(self.cod))]

    def emit(self, b, ins, dims_in, dims_out):
        return emit_spider(b, ins, dims_in, len(self.cod),
                           2 if self.typ == bit else None)


def emit_spider(b: nw.NetworkBuilder, ins, dims_in, n_out: int,
                forced_dim: Optional[int] = None, amplitudes=None) -> List[int]:
    """``tensor.Spider`` with ``EmbeddingTensor`` layers on the wider legs
    (diagram.py:694-707, zw.py:318-364): all legs become ONE shared index; an
    optional amplitude vector is the Z-box diagonal (zw.py:320-329)."""
    if forced_dim is not None:
        spider_dim = forced_dim
    else:
        spider_dim = int(min(dims_in)) if len(ins) > 0 else 2
    legs = []
    for i, d in zip(ins, dims_in):
        if int(d) > spider_dim:
            e = b.new_index(spider_dim)
            b.add(nw.K_EMBED, [i, e], name="Embedding")
            legs.append(e)
        else:
            assert int(d) == spider_dim, "spider leg narrower than the spider"
            legs.append(i)
    if not legs:
        idx = b.new_index(spider_dim)
    else:
        idx = b.merge(legs)
    if amplitudes is not None:
        b.add(nw.K_VEC, [idx], params=np.asarray(amplitudes, dtype=complex), name="Z")
    return [idx] * n_out


class Swap(Box):
    """diagram.py:774-795"""

    def __init__(self, left: Ty, right: Ty):
        super().__init__("Swap", left @ right, right @ left)
        self.left, self.right = left, right
        self.photon_preservation_behaviour = PhotonNumberPreservation.LO

    def conjugate(self):
        return self

    def dagger(self):
        return Swap(self.right, self.left)

    def inflate(self, d):
        return Diagram.swap(Ty._make(self.left.inside * (d if self.left == mode else 1)),
                            Ty._make(self.right.inside * (d if self.right == mode else 1)))

    def determine_output_dimensions(self, input_dims):
        return input_dims[::-1]

    def emit(self, b, ins, dims_in, dims_out):
        return [ins[1], ins[0]]


class Scalar(Box):
    """diagram.py:798-868"""

    def __init__(self, scalar: complex):
        self.scalar = complex(scalar)
        super().__init__("scalar", Mode(0), Mode(0), data=self.scalar)
        self.photon_preservation_behaviour = PhotonNumberPreservation.LO

    def conjugate(self):
        return Scalar(self.scalar.conjugate())

    def dagger(self):
        return Scalar(self.scalar.conjugate())

    def inflate(self, d):
        return self

    def determine_output_dimensions(self, input_dims=None):
        return []

    def emit(self, b, ins, dims_in, dims_out):
        b.mul_scalar(self.scalar)
        return []


class DualRail(Box):
    """diagram.py:871-944"""

    def __init__(self, is_dagger=False, internal_state=None):
        dom = Mode(2) if is_dagger else Bit(1)
        cod = Bit(1) if is_dagger else Mode(2)
        super().__init__("2R", dom, cod)
        self.internal_state = internal_state
        self.is_dagger = is_dagger
        self.photon_preservation_behaviour = PhotonNumberPreservation.CUSTOM

    def photon_number_transform(self, dims_in, dims_out):
        from .zw import Create
        from .zx import Z
        prev = [Z(1, 0) if not self.is_dagger else Z(0, 1),
                Create(1) if not self.is_dagger else Z(1, 0)]
        box = DualRail(self.is_dagger, self.internal_state)
        if not self.is_dagger:
            box.dom = Mode(1)
        prev.append(box)
        return prev if not self.is_dagger else prev[::-1]

    def conjugate(self):
        return self

    def determine_output_dimensions(self, input_dims):
        return [2] if self.is_dagger else [2, 2]

    def emit(self, b, ins, dims_in, dims_out):
        if self.is_dagger:
            out = b.new_index(dims_out[0])
            b.add(nw.K_DUALRAIL, [out, ins[0], ins[1]], name="2R")
            return [out]
        outs = [b.new_index(d) for d in dims_out]
        b.add(nw.K_DUALRAIL, [ins[0]] + outs, name="2R")
        return outs

    def inflate(self, d):
        from .zw import Endo, W
        assert self.internal_state is not None, "Internal state must be provided"
        assert len(self.internal_state) == d, "Internal state must be of len d"
        dgrm = DualRail() >> Diagram.tensor(
            *[(W(d) >> Diagram.tensor(*[Endo(s) for s in self.internal_state]))
              for _ in range(2)])
        return dgrm.dagger() if self.is_dagger else dgrm

    def dagger(self):
        return DualRail(not self.is_dagger, internal_state=self.internal_state)


class PhotonThresholdDetector(Box):
    """diagram.py:947-1002"""

    def __init__(self, is_dagger=False):
        if is_dagger:
            super().__init__("PTD dagger", Bit(1), Mode(1))
        else:
            super().__init__("PTD", Mode(1), Bit(1))
        self.is_dagger = is_dagger
        self.photon_preservation_behaviour = PhotonNumberPreservation.CUSTOM

    def photon_number_transform(self, dims_in, dims_out):
        from .zw import Create, Select
        from .zx import Z
        if self.is_dagger:
            return [Z(1, 0), Create(dims_out[0] - 1)]
        return [Select(0), Z(0, 1)]

    def determine_output_dimensions(self, input_dims):
        if self.is_dagger:
            return [MAX_DIM] * len(input_dims)
        return [2] * len(input_dims)

    def conjugate(self):
        return self

    def dagger(self):
        return PhotonThresholdDetector(not self.is_dagger)

    def inflate(self, d):
        from .zw import Add
        dgrm = Add(d) >> PhotonThresholdDetector()
        return dgrm.dagger() if self.is_dagger else dgrm

    def emit(self, b, ins, dims_in, dims_out):
        out = b.new_index(dims_out[0])
        # canonical [bit, n]
        b.add(nw.K_PTD, [ins[0], out] if self.is_dagger else [out, ins[0]], name="PTD")
        return [out]


class _LayerView:
    """A layer of a replacement *diagram* seen by the light-cone walker as one
    box spanning the whole layer (control.py:125-133 returns a Diagram whose
    layers diagram.py:337-340 iterates)."""

    def __init__(self, dom: Ty, cod: Ty):
        self.dom, self.cod = dom, cod


# ---- light-cone truncation (optyx/utils/utils.py:347-510) ---------------------
# The reference walks ALL earlier layers backwards for every box; the literal
# restatement of that walk lives in the test oracle (oracle/compile.py), which the
# forward formulation below is checked against box by box.
class ConeTracker:
    """Forward formulation of the light-cone walk (utils.py:436-510).

    ``get_max_dim_for_box`` walks ALL earlier layers backwards for every box
    (O(#boxes^2) Python, the dominant host cost of the reference for large or
    batched circuits, SURVEY.md 8(a) row a3').  The same quantity propagates
    forwards: every wire carries the set of photon sources in its past cone -- a
    bitmask over sources, a source being one output of a ``Create`` (weight = its
    photons) or one open mode input of the diagram (weight = 2 * dim).  A box
    hands the union of the cones of its mode-typed inputs to each of its
    outputs; a swap swaps two cones.  The bound of a box is the total weight of
    the union over its mode inputs, + 1, at least 2 -- identical, by
    construction, to what the backward walk returns (checked box by box against
    the oracle's literal walk, tests/test_oracle_compile.py)."""

    def __init__(self, dom: Ty, input_dims: Sequence[int]):
        from .zw import Create, Divide, Endo, Multiply
        self._uncapped = (Swap, Endo, Divide, Multiply)
        self._create = Create
        self.weights: List[int] = []
        self.frontier: List[int] = []
        for t, dim in zip(dom.inside, input_dims):
            if t == "mode":
                self.frontier.append(1 << len(self.weights))
                self.weights.append(2 * int(dim))
            else:
                self.frontier.append(0)
        self._sums = {0: 0}

    def _weight(self, mask: int) -> int:
        s = self._sums.get(mask)
        if s is None:
            s, m, k = 0, mask, 0
            while m:
                if m & 1:
                    s += self.weights[k]
                m >>= 1
                k += 1
            self._sums[mask] = s
        return s

    def max_dim(self, left: int, box) -> float:
        dom = box.dom.inside
        if not dom or isinstance(box, self._uncapped):
            return UNCAPPED
        mask = 0
        frontier = self.frontier
        for k, t in enumerate(dom):
            if t == "mode":
                mask |= frontier[left + k]
        return max(self._weight(mask) + 1, 2)

    def apply(self, left: int, prev_box) -> bool:
        """Push one recorded (replacement) box; False if its span does not fit
        the frontier."""
        Create = self._create
        dom = prev_box.dom.inside
        n_dom, n_cod = len(dom), len(prev_box.cod.inside)
        if left < 0 or left + n_dom > len(self.frontier):
            return False
        if isinstance(prev_box, Swap):
            f = self.frontier
            f[left], f[left + 1] = f[left + 1], f[left]
            return True
        mask = 0
        frontier = self.frontier
        for k, t in enumerate(dom):
            if t == "mode":
                mask |= frontier[left + k]
        if isinstance(prev_box, Create):
            new = []
            for c in range(n_cod):
                new.append(mask | (1 << len(self.weights)))
                self.weights.append(int(prev_box.photons[c]))
        else:
            new = [mask] * n_cod
        self.frontier[left:left + n_dom] = new
        return True


def compile_diagram(d: Diagram, input_dims: Optional[List[int]] = None,
                    builder: Optional[nw.NetworkBuilder] = None,
                    in_inds: Optional[List[int]] = None, nested_eval=None,
                    trace: Optional[list] = None):
    """``Diagram.to_tensor`` (diagram.py:301-375): walk the boxes keeping the
    per-wire truncation dims, cap each box's outputs by its light cone, emit
    leaves.  Returns ``(Network, cod_dims)`` (or ``(out_inds, cod_dims)`` when
    emitting into a caller's builder, used for nested diagrams).  ``trace``
    collects ``(box, dims_in, dims_out)`` per box (the truncation bookkeeping the
    parity tests compare with the oracle's literal walk)."""
    if input_dims is None:
        input_dims = [2 for _ in range(len(d.dom))]
    else:
        assert len(d.dom) == len(input_dims), \
            "Input dims length does not match number of input wires"
    own = builder is None
    b = nw.NetworkBuilder(nested_eval) if own else builder
    layer_dims = [int(x) for x in input_dims]
    wires = [b.new_index(x) for x in layer_dims] if in_inds is None else list(in_inds)
    dom_inds = list(wires)
    cones = ConeTracker(d.dom, input_dims)
    check = __debug__
    for box, left in zip(d.boxes, d.offsets):
        n_dom, n_cod = len(box.dom.inside), len(box.cod.inside)
        end = left + n_dom
        max_dim = cones.max_dim(left, box)
        dims_in = layer_dims[left:end]
        dims_out = [int(max_dim) if i > max_dim else int(i)
                    for i in box.determine_output_dimensions(list(dims_in))]
        repl = box.photon_number_transform(dims_in, dims_out)
        if isinstance(repl, Diagram) and not isinstance(repl, Box):
            repl = [_LayerView(l @ bx.dom @ r, l @ bx.cod @ r) for l, bx, r in repl.layers()]
        for rb in repl:
            if not cones.apply(left, rb):
                # the reference's backward walk clamps such an offset into the frame
                # (utils.py:466-476); no box of the reference records one, so this is a
                # malformed photon_number_transform rather than a case to emulate
                raise NotImplementedError(
                    f"light-cone replacement {rb!r} of {box!r} does not fit the diagram at "
                    f"offset {left}")
        if trace is not None:
            trace.append((box, list(dims_in), list(dims_out)))
        outs = box.emit(b, wires[left:end], dims_in, dims_out)
        if check:
            assert len(outs) == n_cod, (box, outs)
            for o, dd in zip(outs, dims_out):
                assert b.dim(o) == dd, (box, b.dim(o), dd, dims_in, dims_out)
        wires[left:end] = outs
        layer_dims[left:end] = dims_out
    # the per-wire identity Z boxes the reference appends (diagram.py:367-374)
    # are identities on a shared index: nothing to emit
    if own:
        return b.finish(dom_inds, wires), layer_dims
    return wires, layer_dims
