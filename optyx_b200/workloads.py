"""Diagram builders for the BASELINE.json workloads and the reference's own example
circuits; pure host-side constructions on the mirrored optyx
interface, shared by ``bench.py``, ``tools/`` and the tests.  Nothing here touches
the CPU oracle."""
from __future__ import annotations

import random

import numpy as np

from . import classical, photonic, qubits
from .core import channel, diagram, zw, zx  # noqa: F401
from .core.channel import Channel, Discard, Measure, bit, mode, qmode, qubit  # noqa: F401


# ---- reference circuits --------------------------------------------------------
def teleportation():
    """qubits.py:152-160 / test_backends.py:48-56"""
    Id, Z, X, H = qubits.Id, qubits.Z, qubits.X, qubits.H
    return (Id(1) @ Z(0, 2) @ qubits.Scalar(2 ** 0.5) >> Z(1, 2) @ Id(2)
            >> Id(1) @ Z(2, 1) @ Id(1) >> Id(1) @ H() @ Id(1)
            >> qubits.Measure(1) ** 2 @ Id(1)
            >> classical.Id(bit) @ classical.BitControlledGate(X(1, 1, 0.5))
            >> classical.BitControlledGate(Z(1, 1, 0.5)))


def lossy_hom(p=0.8):
    """README.md:176-178: P[(1,)] = 0.2 after AddN(2)"""
    return (photonic.Create(1, 1) >> photonic.PhotonLoss(p) @ photonic.Id(1) >> photonic.BS
            >> photonic.NumberResolvingMeasurement(2) >> classical.AddN(2))


def cfg1_interferometer(width=6, photons=(1, 1, 1, 0, 0, 0), loss=0.9, depth=None, seed=0):
    """lossy 6-mode 3-photon interferometer."""
    depth = width if depth is None else depth
    d = photonic.Create(*photons)
    d = d >> channel.Diagram.tensor(*[photonic.PhotonLoss(loss) for _ in range(width)])
    d = d >> photonic.seeded_ansatz(width, depth, seed)
    return d >> photonic.NumberResolvingMeasurement(width)


def ghz(n, open_qubits=None, noise=None, measure=False):
    """GHZ chain, optional DephasingError, then measure /
    leave ``open_qubits`` open and discard the rest."""
    d = channel.Diagram.tensor(*[qubits.Ket(0) for _ in range(n)])
    d = d >> qubits.H() @ qubits.Id(n - 1)
    for q in range(n - 1):
        d = d >> qubits.Id(q) @ qubits.gate("CX") @ qubits.Id(n - q - 2)
    if noise is not None:
        d = d >> channel.Diagram.tensor(*[qubits.DephasingError(noise) for _ in range(n)])
    if measure:
        return d >> qubits.Measure(n)
    if open_qubits is not None and open_qubits < n:
        d = d >> qubits.Id(open_qubits) @ qubits.Discard(n - open_qubits)
    return d


def random_clifford_t(n, depth, seed=0, close=True):
    """brickwork of {H,S,T} + CX on alternating pairs,
    closed with <0|^n ... |0>^n (pure, scalar amplitude)."""
    rnd = random.Random(seed)
    d = channel.Diagram.tensor(*[qubits.Ket(0) for _ in range(n)])
    for layer in range(depth):
        d = d >> channel.Diagram.tensor(*[qubits.gate(rnd.choice("HST")) for _ in range(n)])
        start = layer % 2
        pairs = (n - start) // 2
        lay = qubits.Id(start)
        for _ in range(pairs):
            lay = lay @ qubits.gate("CX")
        lay = lay @ qubits.Id(n - start - 2 * pairs)
        d = d >> lay
    if close:
        d = d >> channel.Diagram.tensor(*[qubits.Bra(0) for _ in range(n)])
    return d


def distinguishable_bs(state1, state2):
    """test_distinguishability.py:47-94"""
    return (Channel("BS", zw.Create(1, 1, internal_states=(state1, state2))
                    >> photonic.BS.get_kraus())
            >> Measure(qmode ** 2)).inflate(len(state1))


def random_internal_state(rng, dim=2):
    v = rng.random(dim) + 1j * rng.random(dim)
    return v / np.linalg.norm(v)


def cfg3_interferometer(width=4, n_photons=2, measured=2, loss=0.9, seed=0, depth=None,
                        inflate=True):
    """SURVEY 8(d) cfg3 (scaled by arguments): lossy interferometer with
    partially distinguishable photons, ``.inflate(2)``."""
    rng = np.random.default_rng(2025)
    photons = tuple(1 if k < n_photons else 0 for k in range(width))
    states = tuple(random_internal_state(rng) for _ in range(n_photons))
    depth = width if depth is None else depth
    d = photonic.Create(*photons, internal_states=states)
    d = d >> channel.Diagram.tensor(*[photonic.PhotonLoss(loss) for _ in range(width)])
    d = d >> photonic.seeded_ansatz(width, depth, seed)
    d = d >> photonic.NumberResolvingMeasurement(measured) @ photonic.Discard(width - measured) \
        if measured < width else d >> photonic.NumberResolvingMeasurement(width)
    return d.inflate(2) if inflate else d


def fusion_ii_teleportation():
    """test_fusion.py:9-33"""
    correction = classical.BitControlledGate(
        photonic.HadamardBS() >> (photonic.Phase(0.5) @ qmode) >> photonic.HadamardBS())
    channel_bell = qubits.Z(0, 2) @ qubits.Scalar(0.5 ** 0.5) >> \
        photonic.DualRail(1) @ photonic.DualRail(1)
    return (photonic.DualRail(1) @ channel_bell >> photonic.FusionTypeII() @ qmode ** 2
            >> classical.PostselectBit(1) @ correction >> photonic.DualRail(1).dagger())


def fusion_i_spider():
    """test_fusion.py:36-52"""
    correction = classical.BitControlledGate(photonic.Phase(0.5) @ qmode)
    return (photonic.DualRail(1) @ photonic.DualRail(1) >> photonic.FusionTypeI()
            >> qmode ** 2 @ classical.PostselectBit(1) @ bit
            >> qmode @ channel.Swap(qmode, bit) >> channel.Swap(qmode, bit) @ qmode
            >> correction >> photonic.DualRail(1).dagger())


def fusion_link(vector, close="bell"):
    """SURVEY 8(d) cfg5a / README.md:230-280: heralded entanglement link through a
    type-II fusion of two partially distinguishable dual-rail photons."""
    bell_state = qubits.Z(0, 2) @ qubits.Scalar(0.5 ** 0.5)
    enc = photonic.DualRail(1, internal_states=[[1, 0]]) @ \
        photonic.DualRail(1, internal_states=[list(vector)])
    post_select = classical.PostselectBit(1) @ classical.PostselectBit(0)
    experiment = bell_state @ bell_state >> \
        qubits.Id(1) @ (enc >> photonic.FusionTypeII() >> post_select) @ qubits.Id(1)
    tail = bell_state.dagger() if close == "bell" else qubits.Discard(2)
    return (experiment >> tail).inflate(2)


def fusion_teleport_input(phi):
    """fusion-II teleportation with classical feed-forward
    (test_fusion.py:9-33) applied to the input state Ket('+') >> Z(1, 1, phi)."""
    return qubits.Ket("+") >> qubits.Z(1, 1, phi) >> fusion_ii_teleportation()


# ---- named eval cases shared by the parity tests and tests/golden/make_golden.py ------------
EVAL_CASES = {
    "lossy_hom": lambda: lossy_hom(0.8),
    "bs_pure": lambda: photonic.Create(1, 1) >> photonic.BS,
    "mzi_chip_pure": lambda: photonic.Create(1, 1, 1, 1) >> photonic.seeded_ansatz(4, 4, 3),
    "cfg1_4mode": lambda: cfg1_interferometer(4, (1, 1, 0, 0)),
    "cfg1_6mode": lambda: cfg1_interferometer(6, (1, 1, 1, 0, 0, 0)),
    "ghz6_measured": lambda: ghz(6, measure=True),
    "ghz5_noisy_open3": lambda: ghz(5, open_qubits=3, noise=0.95),
    "ghz8_noisy_measured": lambda: ghz(8, noise=0.95, measure=True),
    "bitflip": lambda: qubits.Z(0, 1) >> qubits.BitFlipError(0.43),
    "hom_distinguishable": lambda: distinguishable_bs(np.array([0.6, 0.8j]), np.array([1, 0])),
    "cfg3_small": lambda: cfg3_interferometer(width=3, n_photons=2, measured=2, depth=2),
    "cfg3_4mode": lambda: cfg3_interferometer(width=4, n_photons=2, measured=2),
    "clifford_t_8x6": lambda: random_clifford_t(8, 6, seed=2),
    "fusion_ii": fusion_ii_teleportation,
    "fusion_i": fusion_i_spider,
}


def compile_channel(d):
    """Leaf-descriptor network of a channel diagram as ``B200Backend.eval`` builds it
    (nested sub-networks of controlled boxes are contracted on the device)."""
    from .core.backends import B200Backend
    backend = B200Backend()
    return backend._get_network(d, nested_eval=backend._nested_eval)
