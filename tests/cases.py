"""Shared test circuits: the reference's own known-answer cases (restated from
its tests/docstrings) and the BASELINE.json configs.  ``oracle_eval`` compiles a
channel diagram with the product's host logic and contracts it with the CPU
oracle; the ``-m gpu`` tests run the same diagrams through ``B200Backend``."""
from __future__ import annotations

import random

import numpy as np

from oracle.contract import contract_network
from optyx_b200 import classical, photonic, qubits
from optyx_b200.core import channel, diagram, zw, zx
from optyx_b200.core.channel import Channel, Discard, Measure, bit, mode, qmode, qubit


def oracle_nested(net):
    return contract_network(net)


def compile_channel(d: channel.Diagram):
    """AbstractBackend._get_discopy_tensor (backends.py:406-415)."""
    if d.is_pure:
        return d.get_kraus().to_network(nested_eval=oracle_nested)
    return d.double().to_network(nested_eval=oracle_nested)


def _pad_to(arr: np.ndarray, shape) -> np.ndarray:
    out = np.zeros(shape, dtype=arr.dtype)
    out[tuple(slice(0, s) for s in arr.shape)] = arr
    return out


def oracle_eval(d) -> np.ndarray:
    if hasattr(d, "terms"):
        # formal sum (backends.py:473-476 + diagram.py:737-765): every term on its
        # own, zero-padded to the largest extent of each axis, then added
        arrs = [oracle_eval(t) for t in d.terms]
        shape = tuple(max(a.shape[k] for a in arrs) for k in range(arrs[0].ndim))
        return sum(_pad_to(a, shape) for a in arrs)
    if isinstance(d, channel.Diagram):
        net = compile_channel(d)
    else:
        net = d.to_network(nested_eval=oracle_nested)
    return contract_network(net)


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


def statevector_clifford_t(n, depth, seed=0):
    """Independent check of ``random_clifford_t``: plain gate-by-gate statevector
    simulation (qubit 0 = most significant axis), returns <0..0|C|0..0>."""
    rnd = random.Random(seed)
    psi = np.zeros((2,) * n, dtype=complex)
    psi[(0,) * n] = 1
    G = {"H": np.array([[1, 1], [1, -1]]) / np.sqrt(2), "S": np.diag([1, 1j]),
         "T": np.diag([1, np.exp(1j * np.pi / 4)])}
    for layer in range(depth):
        for q in range(n):
            g = G[rnd.choice("HST")]
            psi = np.moveaxis(np.tensordot(g, psi, axes=([1], [q])), 0, q)
        start = layer % 2
        for p in range((n - start) // 2):
            c, t = start + 2 * p, start + 2 * p + 1
            idx = [slice(None)] * n
            idx[c] = 1
            sub = psi[tuple(idx)]
            psi[tuple(idx)] = np.flip(sub, axis=t - 1 if t > c else t)
    return psi[(0,) * n]


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
