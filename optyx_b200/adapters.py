"""Adapter for an existing Optyx install: quimb tensors -> ``Network``.

``QuimbBackend._process_term`` (optyx/core/backends.py:496-550) turns a
``discopy.tensor.Diagram`` into a quimb ``TensorNetwork`` with ``to_quimb()``
(506), casts integer/bool data to complex (508-511) and contracts it.  A
maintainer who keeps Optyx's own ``double()`` / ``to_tensor()`` can hand the same
tensors to this backend: every quimb tensor becomes one ``K_DENSE`` leaf (its
values travel as host params), index names become integer labels, and
``tn.outer_inds()`` (or the explicit ``output_inds``) gives the result axes.

Only duck typing is used (``.data``, ``.inds``, ``.tensors``, ``.outer_inds()``),
so nothing here imports quimb; tests exercise it with stand-in objects.
"""
from __future__ import annotations

from typing import Dict, Hashable, Iterable, Optional, Sequence

import numpy as np

from .core import network as nw


def network_from_arrays(arrays: Sequence[np.ndarray], inds: Sequence[Sequence[Hashable]],
                        output_inds: Sequence[Hashable], n_dom: int = 0) -> nw.Network:
    """Dense tensors + index names -> ``Network`` (indices may be shared by any
    number of tensors: hyper-indices need no COPY tensor)."""
    label: Dict[Hashable, int] = {}
    dims: Dict[int, int] = {}
    leaves = []
    for arr, names in zip(arrays, inds):
        arr = np.asarray(arr)
        if arr.ndim != len(names):
            raise ValueError(f"tensor of rank {arr.ndim} with {len(names)} index names")
        ids = []
        for name, d in zip(names, arr.shape):
            k = label.setdefault(name, len(label))
            if dims.setdefault(k, int(d)) != int(d):
                raise ValueError(f"index {name!r} has extents {dims[k]} and {d}")
            ids.append(k)
        # int / bool data must become complex (backends.py:508-511)
        vals = np.ascontiguousarray(arr, dtype=np.complex128).reshape(-1)
        leaves.append(nw.Leaf(nw.K_DENSE, tuple(int(d) for d in arr.shape), tuple(ids), 0, vals))
    missing = [q for q in output_inds if q not in label]
    if missing:
        raise ValueError(f"output indices {missing} do not appear on any tensor")
    return nw.Network(leaves, dims, tuple(label[q] for q in output_inds), 1.0 + 0j, n_dom)


def network_from_quimb(tn, output_inds: Optional[Iterable[Hashable]] = None,
                       n_dom: int = 0) -> nw.Network:
    """``tn``: a quimb ``TensorNetwork`` (anything with ``.tensors`` whose items
    have ``.data`` / ``.inds``).  ``output_inds`` defaults to ``tn.outer_inds()``,
    the order ``QuimbBackend`` passes to ``contract(output_inds=...)``
    (backends.py:513-514, 539-545)."""
    tensors = list(tn.tensors)
    out = list(tn.outer_inds()) if output_inds is None else list(output_inds)
    return network_from_arrays([t.data for t in tensors], [tuple(t.inds) for t in tensors],
                               out, n_dom)
