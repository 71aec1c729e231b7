"""Classical-quantum channel diagrams: host-side mirror of
optyx/core/channel.py.

What matters for the hot path is how a circuit becomes ONE single-picture
diagram for the tensor-network compiler:

* ``Diagram.is_pure`` (channel.py:287-329) -> ``get_kraus()`` (331-358), or
* ``Diagram.double()`` (277-285) with ``Channel.double`` (600-641): the CP
  construction -- every quantum wire becomes an adjacent (ket, bra) pair,
  classical wires are copied by spiders, the bra half is ``kraus.conjugate()``
  box by box, environment wires are cupped.
"""
from __future__ import annotations

from typing import List

from . import diagram
from .diagram import Ty as _BaseTy

_CLASSICAL = {"bit": "bit", "mode": "mode", "qubit": "bit", "qmode": "mode"}   # :162-167
_QUANTUM = {"bit": "qubit", "mode": "qmode", "qubit": "qubit", "qmode": "qmode"}  # :168-173


def ob_is_classical(name: str) -> bool:
    return name not in ("qubit", "qmode")                     # channel.py:175-178


class Ty(_BaseTy):
    """channel.py:196-236"""

    def single(self) -> diagram.Ty:
        return diagram.Ty._make(tuple(_CLASSICAL[o] for o in self.inside))

    def double(self) -> diagram.Ty:
        obs: List[str] = []
        for o in self.inside:
            obs.extend([o] if ob_is_classical(o) else [_CLASSICAL[o]] * 2)
        return diagram.Ty._make(tuple(obs))

    @staticmethod
    def from_optyx(ty: diagram.Ty) -> "Ty":
        return Ty._make(tuple(_QUANTUM[o] for o in ty.inside))

    def needs_inflation(self) -> bool:
        return "qmode" in self.inside

    def inflate(self, d: int) -> "Ty":
        obs: List[str] = []
        for o in self.inside:
            obs.extend([o] * d if o == "qmode" else [o])
        return Ty._make(tuple(obs))

    @property
    def is_classical(self) -> bool:
        return all(ob_is_classical(o) for o in self.inside)


bit = Ty("bit")
mode = Ty("mode")
qubit = Ty("qubit")
qmode = Ty("qmode")


_DOUBLE_WIRING: dict = {}


class Diagram(diagram.Diagram):
    """channel.py:245-574"""
    ty_factory = Ty

    @classmethod
    def _swap_box(cls, left, right):
        return Swap(left, right)

    @classmethod
    def _spider_box(cls, n_in, n_out, typ):
        return Spider(n_in, n_out, typ)

    def needs_inflation(self) -> bool:
        return self.dom.needs_inflation() or self.cod.needs_inflation()

    def inflate(self, d: int) -> "Diagram":
        """channel.py:260-275"""
        assert isinstance(d, int) and d > 0
        return self._apply_functor(lambda ty: ty.inflate(d), lambda box: box.inflate(d),
                                   factory=Diagram)

    def double(self) -> diagram.Diagram:
        """channel.py:277-285: functor ob -> ob.double, box -> box.double()."""
        return self._apply_functor(lambda ty: ty.double(), lambda box: box.double(),
                                   factory=diagram.Diagram)

    @property
    def is_pure(self) -> bool:
        """channel.py:287-329 (names kept, including the misleading ones)."""
        are_layers_pure, are_layers_classical = [], []
        for _, generator, _ in self.layers():
            if isinstance(generator, (Discard, Measure)) and \
                    any(not ob_is_classical(o) for o in generator.dom.inside):
                return False
            if hasattr(generator, "env") and generator.env != diagram.Ty():
                return False
            if isinstance(generator, Encode) and \
                    any(ob_is_classical(o) for o in generator.cod.inside):
                return False
            are_layers_pure.append(
                any(ob_is_classical(o) for o in generator.cod.inside)
                or any(ob_is_classical(o) for o in generator.dom.inside)
                or isinstance(generator, Discard))
            are_layers_classical.append(
                all(ob_is_classical(o) for o in generator.cod.inside)
                and all(ob_is_classical(o) for o in generator.dom.inside))
        return not any(are_layers_pure) or all(are_layers_classical)

    def get_kraus(self) -> diagram.Diagram:
        """channel.py:331-358"""
        assert self.is_pure, "Cannot get a Kraus map of non-pure circuit"
        res = diagram.Diagram.id(self.dom.single())
        for left, generator, right in self.layers():
            l, r = diagram.Diagram.id(left.single()), diagram.Diagram.id(right.single())
            if isinstance(generator, Swap):
                mid = diagram.Swap(generator.dom.single()[0], generator.cod.single()[1])
            else:
                mid = generator.kraus
            res = res >> (l @ mid @ r)
        return res

    @classmethod
    def from_bosonic_operator(cls, n_modes: int, operators, scalar=1):
        """channel.py:459-479: a product of creation / annihilation operators
        ``[(mode, is_dagger), ...]`` as a ZW channel (annihilation = split one photon
        off and select it)."""
        from . import zw
        d = Diagram.id(qmode ** n_modes)
        annil = Channel("annil", zw.Split(2) >> zw.Select(1) @ zw.Id(1))
        create = annil.dagger()
        for idx, dagger in operators:
            if not 0 <= idx < n_modes:
                raise ValueError(f"Index {idx} out of bounds.")
            box = create if dagger else annil
            d = d >> (Diagram.id(qmode ** idx) @ box @ Diagram.id(qmode ** (n_modes - idx - 1)))
        if scalar != 1:
            d = Channel(f"{scalar}", diagram.Scalar(scalar)) @ d
        return d

    def eval(self, backend=None, **kwargs):
        """channel.py:564-574; the default backend here is the B200 one."""
        from .backends import B200Backend
        if backend is None:
            backend = B200Backend()
        return backend.eval(self, **kwargs)


Diagram.diagram_factory = Diagram


class Sum(diagram.Sum):
    """Formal sum of channel diagrams (channel.py:697-722)."""

    def double(self) -> diagram.Sum:
        return diagram.Sum([t.double() for t in self.terms])            # channel.py:706-707

    def get_kraus(self):
        if not self.terms:
            return diagram.Scalar(0)
        return diagram.Sum([t.get_kraus() for t in self.terms])         # channel.py:715-722

    @property
    def is_pure(self) -> bool:
        return all(t.is_pure for t in self.terms)

    def eval(self, backend=None, **kwargs):
        from .backends import B200Backend
        if backend is None:
            backend = B200Backend()
        return backend.eval(self, **kwargs)


Diagram.sum_factory = Sum


class _CBox(Diagram):
    """Common base of channel-level generators."""

    def __init__(self, name: str, dom: Ty, cod: Ty):
        self.name = name
        Diagram.__init__(self, dom, cod, [self], [0])

    def __repr__(self):
        return self.name


class Channel(_CBox):
    """Channel initialised by its Kraus map, channel.py:577-681."""

    def __init__(self, name, kraus, dom=None, cod=None, env=None):
        env = diagram.Ty() if env is None else env
        assert isinstance(kraus, diagram.Diagram)
        if dom is None:
            dom = Ty.from_optyx(kraus.dom)
        if cod is None:
            cod = Ty.from_optyx(kraus.cod)
        assert kraus.dom == dom.single(), (kraus.dom, dom)
        assert kraus.cod == cod.single() @ env, (kraus.cod, cod, env)
        self.kraus, self.env = kraus, env
        super().__init__(name, dom, cod)

    def double(self) -> diagram.Diagram:
        """The CP construction, channel.py:600-641 (memoised per instance: boxes are
        immutable values and a circuit re-uses the same gate objects many times)."""
        hit = self.__dict__.get("_doubled")
        if hit is None:
            hit = self.__dict__["_doubled"] = self._double()
        return hit

    def _double(self) -> diagram.Diagram:
        D = diagram.Diagram

        def get_spiders(ty: Ty) -> diagram.Diagram:
            spiders = D.id(diagram.Ty())
            for o in ty.inside:
                if ob_is_classical(o):
                    box = diagram.Spider(1, 2, diagram.Ty(_CLASSICAL[o]))
                else:
                    box = D.id(diagram.Ty(_CLASSICAL[o]) ** 2)
                spiders = spiders @ box
            return spiders

        def get_perm(n):
            return sorted(sorted(list(range(n))), key=lambda i: i % 2)

        # the wiring around ``kraus (x) conj(kraus)`` depends on the types only:
        # built once per (dom, cod, env) -- diagrams are immutable values
        key = (self.dom.inside, self.cod.inside, self.env.inside)
        hit = _DOUBLE_WIRING.get(key)
        if hit is None:
            cod = self.cod.single()
            top_spiders = get_spiders(self.dom)
            top_perm = D.permutation(get_perm(len(top_spiders.cod)), top_spiders.cod)
            swap_env = D.id(cod @ self.env) @ D.swap(cod, self.env)
            discard = D.id(cod) @ D.spiders(2, 0, self.env) @ D.id(cod)
            new_cod = diagram.Ty._make(sum(((o, o) for o in cod.inside), ()))
            bot_perm = D.permutation(get_perm(2 * len(cod)), new_cod).dagger()
            bot_spiders = get_spiders(self.cod).dagger()
            hit = (top_spiders >> top_perm, swap_env >> discard >> bot_perm >> bot_spiders)
            if len(_DOUBLE_WIRING) < 4096:
                _DOUBLE_WIRING[key] = hit
        top, bot = hit
        return top >> (self.kraus @ self.kraus.conjugate()) >> bot

    def dagger(self):
        return Channel(name=self.name + ".dagger()", kraus=self.kraus.dagger(),
                       dom=self.cod, cod=self.dom)

    def inflate(self, d):
        return Channel(name=self.name + f"^{d}",
                       kraus=self.kraus.inflate(d) if self.needs_inflation() else self.kraus,
                       dom=self.dom.inflate(d), cod=self.cod.inflate(d))


class Spider(Channel):
    """channel.py:684-696"""

    def __init__(self, n_legs_in: int, n_legs_out: int, typ: Ty, **params):
        kraus = diagram.Spider(n_legs_in, n_legs_out, typ.single())
        super().__init__(f"Spider({n_legs_in}, {n_legs_out})", kraus,
                         typ ** n_legs_in, typ ** n_legs_out)
        self.typ = typ

    def dagger(self):
        return Spider(len(self.cod), len(self.dom), self.typ)


class Swap(_CBox):
    """channel.py:771-773; doubled as a block swap of the doubled types."""

    def __init__(self, left: Ty, right: Ty):
        self.left, self.right = left, right
        super().__init__("Swap", left @ right, right @ left)

    @property
    def kraus(self):
        return diagram.Diagram.swap(self.left.single(), self.right.single())

    def double(self):
        return diagram.Diagram.swap(self.left.double(), self.right.double())

    def dagger(self):
        return Swap(self.right, self.left)

    def inflate(self, d):
        return Diagram.swap(self.left.inflate(d), self.right.inflate(d))


class CQMap(_CBox):
    """Channel initialised by its doubled map, channel.py:724-768."""

    def __init__(self, name, density_matrix, dom, cod):
        assert isinstance(density_matrix, diagram.Diagram)
        assert density_matrix.dom == dom.double(), (density_matrix.dom, dom.double())
        assert density_matrix.cod == cod.double(), (density_matrix.cod, cod.double())
        self.density_matrix = density_matrix
        super().__init__(name, dom, cod)

    @property
    def kraus(self):
        # classical maps are "pure" (channel.py:323-329): their single picture
        # is the density-matrix diagram itself
        return self.density_matrix

    def double(self):
        return self.density_matrix

    def dagger(self):
        return CQMap(name=self.name + ".dagger()",
                     density_matrix=self.density_matrix.dagger(), dom=self.cod, cod=self.dom)

    def inflate(self, d):
        return CQMap(name=self.name + f"^{d}",
                     density_matrix=self.density_matrix.inflate(d) if self.needs_inflation()
                     else self.density_matrix,
                     dom=self.dom.inflate(d), cod=self.cod.inflate(d))


class Measure(Channel):
    """channel.py:776-811"""

    def __init__(self, dom: Ty):
        cod = Ty._make(tuple(_CLASSICAL[o] for o in dom.inside))
        super().__init__(name="Measure", kraus=diagram.Diagram.id(dom.single()), dom=dom, cod=cod)

    def inflate(self, d):
        from .zw import Add
        res = Diagram.id(Ty())
        for ob in self.dom:
            if ob.needs_inflation():
                res = res @ (Measure(ob ** d) >> CQMap("Gather photons", Add(d), mode ** d, mode))
            else:
                res = res @ Measure(ob)
        return res


class Encode(Channel):
    """channel.py:814-883"""

    def __init__(self, dom: Ty, internal_states=None):
        cod = Ty._make(tuple(_QUANTUM[o] for o in dom.inside))
        if internal_states is not None:
            if not isinstance(internal_states, tuple):
                internal_states = (internal_states,)
            assert len(internal_states) == sum(1 for o in dom.inside if o == "mode"), \
                "# of internal states must match the number of modes in dom"
            assert len(set(len(i) for i in internal_states)) == 1
        super().__init__(name="Encode", kraus=diagram.Diagram.id(dom.single()), dom=dom, cod=cod)
        self.internal_states = internal_states

    def inflate(self, d):
        from .zw import Add, Endo
        if any(o == "mode" for o in self.dom.inside):
            assert self.internal_states is not None, \
                "Internal states must be provided for encoding"
            assert all(len(s) == d for s in self.internal_states)
        amps_iter = iter(self.internal_states or [])
        res = Diagram.id(Ty())
        for ob in self.dom:
            if ob == mode:
                amps = next(amps_iter)
                amp_layer = diagram.Diagram.tensor(*[Endo(a) for a in amps])
                res = res @ (CQMap("Add†", Add(d).dagger(), mode, mode ** d)
                             >> Encode(mode ** d) >> Channel("Amplitudes", amp_layer))
            elif ob == qmode:
                res = res @ Encode(qmode ** d)
            else:
                res = res @ Encode(ob)
        return res


class Discard(Channel):
    """channel.py:886-903"""

    def __init__(self, dom: Ty):
        env = dom.single()
        super().__init__("Discard", diagram.Diagram.id(dom.single()), dom=dom, cod=Ty(), env=env)

    def inflate(self, d):
        return Discard(self.dom.inflate(d))
