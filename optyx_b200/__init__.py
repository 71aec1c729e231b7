"""optyx_b200 -- a B200-native evaluation backend for Optyx diagrams.

The hot path ``diagram.eval(backend)`` of quantinuum-dev/optyx (dense leaf
arrays -> pairwise contractions -> sliced sum -> prob_dist) in hand-written
sm_100a CUDA behind a C ABI, with the host-side mirror of the reference's
diagram / channel / backend interface needed to drive it.
"""
from .core.channel import bit, mode, qubit, qmode, Channel, Diagram  # noqa: F401
from .core.backends import B200Backend, EvalResult, PermanentBackend, StateType  # noqa: F401

__all__ = ["bit", "mode", "qubit", "qmode", "Channel", "Diagram", "B200Backend",
           "PermanentBackend", "EvalResult", "StateType"]
