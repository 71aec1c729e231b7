"""Structural fingerprint of a diagram: equal fingerprints <=> same boxes, same
parameters, same wiring, hence the same compiled network.

The reference re-runs ``double()`` + ``to_tensor()`` (channel.py:277-285,
diagram.py:301-375) on every ``eval``; only its cotengra ``Reusable*Optimizer`` keeps
anything between calls (the contraction path, backends.py:49-56).  Here an identical
diagram that was re-built from scratch (the usual pattern: a function returning the
circuit) skips the host compile altogether.  The fingerprint reads EVERY instance
attribute of every box (``vars``), so a parameter cannot be missed; a value it does
not understand makes the diagram uncacheable (``None``), never wrongly cached."""
from __future__ import annotations

import enum
from typing import Optional

import numpy as np


class _Opaque(Exception):
    pass


def fingerprint(diagram) -> Optional[tuple]:
    """Hashable key, or ``None`` when the diagram holds something that cannot be
    compared by value.  Callables are compared by identity and kept alive by the key."""
    try:
        keep: list = []
        key = _diagram(diagram, {}, keep)
        return (key, _KeepAlive(keep)) if keep else (key, None)
    except (_Opaque, RecursionError):
        return None


class _KeepAlive:
    """Holds the callables whose ``id`` is part of a key (an id is only unique while the
    object lives); all instances compare equal so that it does not affect the key."""
    __slots__ = ("objs",)

    def __init__(self, objs):
        self.objs = objs

    def __eq__(self, other):
        return isinstance(other, _KeepAlive)

    def __hash__(self):
        return 0


def _diagram(d, memo, keep):
    boxes = d.boxes
    if len(boxes) == 1 and boxes[0] is d:
        return _box(d, memo, keep)
    return ("D", type(d).__qualname__, d.dom.inside, d.cod.inside, tuple(d.offsets),
            tuple(_box(b, memo, keep) for b in boxes))


def _box(box, memo, keep):
    hit = memo.get(id(box))
    if hit is not None:
        return hit
    memo[id(box)] = ("cycle", len(memo))
    items = []
    for k, v in vars(box).items():
        if k in ("boxes", "offsets", "_doubled"):   # a box lists itself; memoised derived data
            continue
        items.append((k, _value(v, memo, keep)))
    res = ("B", type(box).__module__, type(box).__qualname__, tuple(items))
    memo[id(box)] = res
    return res


def _value(v, memo, keep):
    if v is None or isinstance(v, (bool, int, str, bytes)):
        return v
    if isinstance(v, (float, complex)):
        return (type(v).__name__, repr(v))          # repr: -0.0, nan and 1 vs 1.0 stay distinct
    if isinstance(v, np.generic):
        return ("np", v.dtype.str, repr(v.item()))
    if isinstance(v, np.ndarray):
        if v.dtype == object:
            return ("ao", v.shape, tuple(_value(x, memo, keep) for x in v.reshape(-1)))
        return ("a", v.shape, v.dtype.str, v.tobytes())
    if isinstance(v, (tuple, list)):
        return (type(v).__name__, tuple(_value(x, memo, keep) for x in v))
    if isinstance(v, dict):
        return ("dict", tuple(sorted((repr(k), _value(x, memo, keep)) for k, x in v.items())))
    if isinstance(v, enum.Enum):
        return ("enum", type(v).__qualname__, v.name)
    if hasattr(v, "boxes") and hasattr(v, "offsets") and hasattr(v, "dom"):
        return _diagram(v, memo, keep)
    if hasattr(v, "inside") and isinstance(getattr(v, "inside"), tuple):
        return ("ty", v.inside)
    if callable(v) and not hasattr(v, "__dict__") or isinstance(v, type(lambda: 0)):
        keep.append(v)
        return ("fn", id(v))
    if hasattr(v, "__dict__"):                       # IndexableAmplitudes and the like
        return ("o", type(v).__qualname__,
                tuple((k, _value(x, memo, keep)) for k, x in vars(v).items()))
    raise _Opaque(type(v).__name__)
