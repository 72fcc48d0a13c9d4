"""Problem definitions for the measurement scripts, built with the PRODUCT's own models (casq_b200.models) — the oracle
is test infrastructure and is not imported by anything under scripts/.  Constants: ibmq_manila (reference
src/casq/models/transmon_model.py:54-115) and the carrier frequencies of docs/source/user_guide/pulse_backend.rst:86-99
(u5..u7 are absent from the docs example: the u-channel LO is the target qubit's drive LO)."""
import numpy as np

from casq_b200._native import NativeModel
from casq_b200.models import HamiltonianModel, TransmonModel

DT = 2.2222222222222221e-10
CARRIERS = {
    "d0": 4962770879.920025, "d1": 4838412258.764764, "d2": 5036989248.286842, "d3": 4951300212.210368, "d4": 5066350584.469812,
    "u0": 4838412258.764764, "u1": 4962770879.920025, "u2": 5036989248.286842, "u3": 4838412258.764764, "u4": 4951300212.210368,
    "u5": 5036989248.286842, "u6": 5066350584.469812, "u7": 4951300212.210368,
}


def manila(extracted_qubits=None):
    """(static [d,d], operators [K,d,d], channels) of the MANILA dictionary restricted to `extracted_qubits` (None = all 5)."""
    h = TransmonModel.manila(extracted_qubits=extracted_qubits)
    return np.asarray(h.static_operator), np.asarray(h.operators), list(h.channels)


def chain(nq):
    """The first `nq` transmons of ibmq_manila with their exchange couplings, drive channels d_q and control channels u."""
    v = TransmonModel.MANILA["vars"]
    qm = {q: TransmonModel.TransmonProperties(frequency=v[f"wq{q}"], anharmonicity=v[f"delta{q}"], drive=v[f"omegad{q}"])
          for q in range(nq)}
    cm = {(q, q + 1): v[f"jq{q}q{q + 1}"] for q in range(nq - 1)}
    h = TransmonModel(qubit_map=qm, coupling_map=cm)
    return np.asarray(h.static_operator), np.asarray(h.operators), list(h.channels)


def native(static, operators, channels, frame="full", rwa=None, device=0):
    """NativeModel for (static, operators, channels); frame = "full" (H0), "diag" (diag(H0)) or None."""
    rf = {"full": static, "diag": np.diag(static).real, None: None}[frame]
    return NativeModel(static, operators, [CARRIERS[c] for c in channels], DT, rotating_frame=rf, rwa_cutoff_freq=rwa, device=device)


def t1t2(nq, T1=100e-6, T2=80e-6):
    """Per qubit L1 = sqrt(1/T1) a, Lphi = sqrt(2 gamma_phi) a^+ a with gamma_phi = 1/T2 - 1/(2 T1) (SURVEY 8(d) cfg 4)."""
    a = np.diag(np.sqrt(np.arange(1, 3)), k=1).astype(complex)
    n = a.conj().T @ a
    out = []
    for q in range(nq):
        def emb(m_):
            full = np.array([[1.0 + 0j]])
            for s in range(nq):
                full = np.kron(m_ if s == q else np.eye(3), full)
            return full
        out += [np.sqrt(1 / T1) * emb(a), np.sqrt(2 * (1 / T2 - 1 / (2 * T1))) * emb(n)]
    return np.array(out)
