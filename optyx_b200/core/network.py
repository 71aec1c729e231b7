"""Tensor-network IR handed from the host-side diagram compiler to the
CUDA executor (and, as plain data, to the CPU oracle in ``oracle/``).

The reference goes ``diagram.Diagram.to_tensor`` -> ``discopy.tensor.Diagram``
-> ``.to_quimb()`` -> quimb ``TensorNetwork`` (optyx/core/diagram.py:301-375,
optyx/core/backends.py:496-550).  Here the same information is a flat list of
*leaf descriptors* (what dense truncated array to build, reference
``Box.truncation``, diagram.py:536-590) plus integer index labels.  COPY
tensors, swaps and identities never become leaves: a spider is a shared
(hyper-)index, a swap is a relabel.

A leaf descriptor is deliberately tiny and closed-form so the array can be
built *on the device* by ``b2_build_leaves`` (csrc/leaf_build.cu).  Values that
the reference computes with arbitrary host Python (``s ** k`` for ``Endo``,
user arrays, classical-function truth tables) travel in ``params`` unchanged.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import math

import numpy as np

# ---- leaf kinds (mirrored in include/optyx_b200.h and oracle/leaves.py) ----
K_ONEHOT = 1    # rank 1: e_pos (all zero if pos >= d)      Create/Select zw.py:450-460, 567-573
K_ONES = 2      # rank 1: all ones (free end of a spider)
K_VEC = 3       # rank 1: host values (diagonal of a Z box)  zw.py:320-329, 641-646
K_DENSE = 4     # any rank: host values, C order              diagram.py:542-548, zx.py:304-309
K_SUM = 5       # [n, n_1..n_k]: sqrt(multinomial) iff sum n_i == n   W  zw.py:231-244
K_ADD = 6       # [s, n_1..n_k]: 1 iff sum n_i == s                   Add zw.py:676-683
K_MUL = 7       # [p, a, b]: 1 iff a*b == p                           Multiply zw.py:720-727
K_DIV = 8       # [q, a, b]: 1 iff b != 0 and a == q*b                Divide zw.py:763-772
K_MOD2 = 9      # [b, n]: 1 iff n % 2 == b                            Mod2 zw.py:818-825
K_DUALRAIL = 10  # [b, m0, m1]: 1 iff (m0, m1) == (1-b, b)            diagram.py:904-913
K_PTD = 11      # [b, n]: 1 iff b == min(n, 1)                        diagram.py:975-983
K_EMBED = 12    # [i, j]: 1 iff i == j (rectangular identity)         diagram.py:1005-1027

KIND_NAMES = {
    K_ONEHOT: "onehot", K_ONES: "ones", K_VEC: "vec", K_DENSE: "dense",
    K_SUM: "w", K_ADD: "add", K_MUL: "mul", K_DIV: "div", K_MOD2: "mod2",
    K_DUALRAIL: "dualrail", K_PTD: "ptd", K_EMBED: "embed",
}

MAX_LEAF_RANK = 8       # closed-form leaves decoded per axis on the device (b2_leaf_desc.dims)


def closed_form_values(kind: int, shape: Tuple[int, ...], ipar: int = 0) -> np.ndarray:
    """Host values (C order, complex128) of a closed-form leaf, used ONLY when its rank
    exceeds what the device descriptor decodes (``MAX_LEAF_RANK``: a ``W`` / ``Add`` with
    eight or more legs); such a leaf travels as ``K_DENSE`` like every other host-valued
    array.  Same predicates as csrc/leaf_build.cu."""
    shape = tuple(int(d) for d in shape)
    idx = np.indices(shape, dtype=np.int64) if shape else np.zeros((0,), dtype=np.int64)
    if kind == K_ONES:
        return np.ones(shape, dtype=np.complex128).reshape(-1)
    if kind == K_ONEHOT:
        return (idx[0] == ipar).astype(np.complex128).reshape(-1)
    if kind in (K_SUM, K_ADD):
        rest = idx[1:].sum(axis=0)
        hit = rest == idx[0]
        if kind == K_ADD:
            return hit.astype(np.complex128).reshape(-1)
        fact = np.array([math.factorial(k) for k in range(int(max(shape)) * len(shape) + 1)],
                        dtype=object)
        num = fact[rest]
        den = np.ones(shape, dtype=object)
        for a in range(1, len(shape)):
            den = den * fact[idx[a]]
        vals = np.zeros(shape, dtype=np.float64)
        vals[hit] = [math.sqrt(int(n) // int(d)) for n, d in zip(num[hit], den[hit])]
        return vals.astype(np.complex128).reshape(-1)
    if kind == K_MUL:
        return (np.prod(idx[1:], axis=0) == idx[0]).astype(np.complex128).reshape(-1)
    if kind == K_DIV:
        return ((idx[2] != 0) & (idx[1] == idx[0] * idx[2])).astype(np.complex128).reshape(-1)
    if kind == K_MOD2:
        return (idx[1] % 2 == idx[0]).astype(np.complex128).reshape(-1)
    if kind == K_DUALRAIL:
        return ((idx[1] == 1 - idx[0]) & (idx[2] == idx[0])).astype(np.complex128).reshape(-1)
    if kind == K_PTD:
        return (idx[0] == np.minimum(idx[1], 1)).astype(np.complex128).reshape(-1)
    if kind == K_EMBED:
        return (idx[0] == idx[1]).astype(np.complex128).reshape(-1)
    raise ValueError(f"leaf kind {kind} has no closed form")


@dataclass
class Leaf:
    """One dense leaf: ``kind`` + ``shape`` say what to build, ``inds`` how
    it is wired.  ``ipar`` is an integer parameter (one-hot position);
    ``params`` complex values for VEC/DENSE kinds (C order): shape ``(size,)``, or
    ``(size, batch)`` when the diagram was built with ARRAY-valued parameters (one
    compile for a whole batch of evals: column e belongs to eval e)."""
    kind: int
    shape: Tuple[int, ...]
    inds: Tuple[int, ...]
    ipar: int = 0
    params: np.ndarray | None = None
    name: str = ""

    @property
    def size(self) -> int:
        return math.prod(self.shape)


@dataclass
class Network:
    """Leaves + wiring.  ``dims[i]`` is the extent of index ``i``;
    ``output_inds`` are the open indices in result-axis order (``dom + cod``
    wires of the tensor diagram, backends.py:485-491) and may repeat an index
    (e.g. a bare ``Spider(1, 2)``).  ``scalar`` multiplies the result
    (product of all ``Scalar`` boxes and closed loops)."""
    leaves: List[Leaf] = field(default_factory=list)
    dims: Dict[int, int] = field(default_factory=dict)
    output_inds: Tuple[int, ...] = ()
    scalar: complex = 1.0 + 0j
    n_dom: int = 0          # how many of output_inds belong to the domain

    @property
    def output_shape(self) -> Tuple[int, ...]:
        return tuple(self.dims[i] for i in self.output_inds)

    @property
    def batch(self):
        """Number of evals this network stands for (array-valued parameters), or None."""
        sizes = {l.params.shape[1] for l in self.leaves
                 if l.params is not None and l.params.ndim == 2}
        if np.ndim(self.scalar) > 0 and np.size(self.scalar) > 1:
            sizes.add(int(np.size(self.scalar)))
        if not sizes:
            return None
        if len(sizes) > 1:
            raise ValueError(f"array-valued parameters of different lengths: {sorted(sizes)}")
        return sizes.pop()

    def params_matrix(self, batch: int) -> np.ndarray:
        """Host values of every VEC / DENSE leaf for ``batch`` evals: ``[batch, P]``."""
        cols = []
        for l in self.leaves:
            if l.kind in (K_VEC, K_DENSE):
                p = l.params
                cols.append(np.broadcast_to(p[:, None], (l.size, batch)) if p.ndim == 1 else p)
        if not cols:
            return np.zeros((batch, 1), dtype=np.complex128)
        return np.ascontiguousarray(np.concatenate(cols, axis=0).T)

    def structure_key(self) -> tuple:
        """Hashable description of everything except VEC/DENSE values and the
        scalar: two networks with equal keys share a contraction plan."""
        return (
            tuple((l.kind, l.shape, l.inds, l.ipar) for l in self.leaves),
            self.output_inds,
        )

    def arena_elems(self) -> int:
        return sum(l.size for l in self.leaves)

    def validate(self) -> None:
        """Raises ``ValueError`` on a malformed network (also under ``python -O``)."""
        for leaf in self.leaves:
            if len(leaf.shape) != len(leaf.inds):
                raise ValueError(f"leaf {leaf.name!r}: {len(leaf.shape)} axes, {len(leaf.inds)} indices")
            # host-valued leaves (VEC / DENSE) are flat on the device, so their rank is
            # unbounded; closed-form kinds are decoded per axis by the array-build kernel
            if leaf.kind not in (K_VEC, K_DENSE) and len(leaf.shape) > MAX_LEAF_RANK:
                raise ValueError(f"closed-form leaf {leaf.name!r} of rank {len(leaf.shape)} > "
                                 f"{MAX_LEAF_RANK}: NetworkBuilder.add converts these to DENSE")
            for d, i in zip(leaf.shape, leaf.inds):
                if self.dims.get(i) != d:
                    raise ValueError(f"leaf {leaf.name!r}: index {i} has extent {self.dims.get(i)}, "
                                     f"axis has {d}")
            if leaf.kind in (K_VEC, K_DENSE):
                if leaf.params is None or leaf.params.shape[0] != leaf.size or leaf.params.ndim > 2:
                    raise ValueError(f"leaf {leaf.name!r}: host values missing or of the wrong size")
        for i in self.output_inds:
            if i not in self.dims:
                raise ValueError(f"open index {i} has no extent")


class NetworkBuilder:
    """Accumulates leaves while the diagram compiler walks the boxes.

    Wires carry integer index labels.  ``merge`` implements a spider
    (``tensor.Spider`` emitted at diagram.py:689, 705-707; zw.py:357-359) by
    union-find, so COPY tensors are never materialised."""

    def __init__(self, nested_eval=None):
        self._parent: List[int] = []
        self._dim: List[int] = []
        self.leaves: List[Leaf] = []
        self.scalar: complex = 1.0 + 0j
        # callable(Network) -> ndarray used by leaves that are themselves
        # contracted sub-networks (BitControlledBox, control.py:164-191)
        self.nested_eval = nested_eval

    # -- indices ------------------------------------------------------------
    def new_index(self, dim: int) -> int:
        self._parent.append(len(self._parent))
        self._dim.append(int(dim))
        return len(self._parent) - 1

    def find(self, i: int) -> int:
        root = i
        while self._parent[root] != root:
            root = self._parent[root]
        while self._parent[i] != root:
            self._parent[i], i = root, self._parent[i]
        return root

    def dim(self, i: int) -> int:
        return self._dim[self.find(i)]

    def merge(self, inds: Sequence[int]) -> int:
        """Identify all ``inds`` (they must have equal extent)."""
        roots = [self.find(i) for i in inds]
        d = self._dim[roots[0]]
        for r in roots[1:]:
            assert self._dim[r] == d, "spider legs must have equal dimension"
            if r != roots[0]:
                self._parent[r] = roots[0]
        return roots[0]

    # -- leaves -------------------------------------------------------------
    def add(self, kind: int, inds: Sequence[int], ipar: int = 0,
            params=None, name: str = "") -> None:
        shape = tuple(self.dim(i) for i in inds)
        if params is None and kind not in (K_VEC, K_DENSE) and len(shape) > MAX_LEAF_RANK:
            kind, params, ipar = K_DENSE, closed_form_values(kind, shape, ipar), 0
        if params is not None:
            params = np.ascontiguousarray(params, dtype=np.complex128)
            size = math.prod(shape)
            if params.size == size:
                params = params.reshape(-1)
            elif params.ndim >= 1 and params.size == size * params.shape[-1]:
                # array-valued parameters: the last axis runs over the evals of a batch
                params = params.reshape(size, params.shape[-1])
            else:
                raise ValueError(f"leaf {name!r}: {params.size} host values for {size} elements")
        self.leaves.append(Leaf(kind, shape, tuple(inds), ipar, params, name))

    def mul_scalar(self, value) -> None:
        if np.ndim(value) > 0 and np.size(value) > 1:
            self.scalar = self.scalar * np.asarray(value, dtype=np.complex128).reshape(-1)
        else:
            self.scalar = self.scalar * complex(np.asarray(value).reshape(-1)[0]
                                                if np.ndim(value) > 0 else value)

    # -- finish ---------------------------------------------------------------
    def finish(self, dom_inds: Sequence[int], cod_inds: Sequence[int]) -> Network:
        find = self.find
        out = tuple(find(i) for i in list(dom_inds) + list(cod_inds))
        # compact relabel in first-use order (leaves, then open indices)
        relabel: Dict[int, int] = {}
        leaves: List[Leaf] = []
        for leaf in self.leaves:
            inds = []
            for i in leaf.inds:
                r = find(i)
                k = relabel.get(r)
                if k is None:
                    k = relabel[r] = len(relabel)
                inds.append(k)
            leaves.append(Leaf(leaf.kind, leaf.shape, tuple(inds), leaf.ipar, leaf.params, leaf.name))
        scalar = self.scalar
        # a closed index touched by no leaf is a loop: Spider(0,0,Dim(d)) = d
        # (checked before the open ends below add their ``ones`` leaves)
        open_roots = set(out)
        for i in range(len(self._parent)):
            if self._parent[i] == i and i not in relabel and i not in open_roots:
                scalar = scalar * self._dim[i]
        # an open index touched by no leaf is the free end of a spider: ones
        for i in dict.fromkeys(out):
            if i not in relabel:
                relabel[i] = len(relabel)
                leaves.append(Leaf(K_ONES, (self._dim[i],), (relabel[i],), 0, None, "ones"))
        net = Network(
            leaves=leaves,
            dims={k: self._dim[i] for i, k in relabel.items()},
            output_inds=tuple(relabel[i] for i in out),
            scalar=scalar,
            n_dom=len(dom_inds),
        )
        net.validate()
        return net


def pad_outputs_to_common_shape(nets: Sequence[Network]) -> List[Network]:
    """Terms of a formal sum: every output (cod) wire is embedded into the per-wire
    maximum dimension over the terms with an ``EmbeddingTensor`` leaf
    (optyx/core/diagram.py:739-757, 1005-1027), so that the results can be added.
    Domain axes are fixed by ``input_dims`` and must already agree."""
    nets = list(nets)
    if not nets:
        return nets
    shapes = [n.output_shape for n in nets]
    rank = len(shapes[0])
    if any(len(s) != rank for s in shapes):
        raise ValueError("terms of a sum have different numbers of open wires")
    target = tuple(max(s[p] for s in shapes) for p in range(rank))
    out = []
    for net in nets:
        n_dom = net.n_dom
        if any(net.output_shape[p] != target[p] for p in range(n_dom)):
            raise ValueError("terms of a sum disagree on the input dimensions")
        leaves, dims = list(net.leaves), dict(net.dims)
        outs = list(net.output_inds)
        nxt = max(dims, default=-1) + 1
        for p in range(n_dom, rank):
            d = net.output_shape[p]
            if d == target[p]:
                continue
            dims[nxt] = target[p]
            leaves.append(Leaf(K_EMBED, (d, target[p]), (outs[p], nxt), 0, None, "embed"))
            outs[p] = nxt
            nxt += 1
        out.append(Network(leaves, dims, tuple(outs), net.scalar, n_dom))
    return out
