"""Oracle: Hamiltonian string parser, rotating frame, RWA and generator evaluation.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED: restates qiskit-dynamics 0.4.1
semantics reached from casq at
  * src/casq/models/hamiltonian_model.py:85-97   (parse_backend_hamiltonian_dict, frame default)
  * src/casq/models/transmon_model.py:54-115,141-161 (h_str vocabulary and constants)
  * src/casq/backends/qiskit/qiskit_pulse_backend.py:174-183 (Solver ctor arguments)
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field

import numpy as np

# --------------------------------------------------------------------------------------
# Numeric input fixtures (the only on-disk numbers the reference holds for this path):
# src/casq/models/transmon_model.py:94-114 and docs/source/user_guide/pulse_backend.rst:86-99
# --------------------------------------------------------------------------------------
_WQ = [31179079102.853794, 30397743782.610542, 31649945798.50227, 31107813662.24873, 31825180853.3539]
_DELTA = [-2165345334.8252344, -2169482392.6367006, -2152313197.3287387, -2158766696.6684937, -2149525690.7311115]
_OMEGAD = [926545606.6640488, 892870223.8110852, 927794953.0001632, 921439621.8693779, 1150709205.1097605]
_J = {(0, 1): 11845444.218797993, (1, 2): 11967839.68906386, (2, 3): 12402113.956012368, (3, 4): 12186910.37040823}

MANILA_DT = 2.2222222222222221e-10
MANILA_CARRIERS = {
    "d0": 4962770879.920025, "d1": 4838412258.764764, "d2": 5036989248.286842,
    "d3": 4951300212.210368, "d4": 5066350584.469812,
    "u0": 4838412258.764764, "u1": 4962770879.920025, "u2": 5036989248.286842,
    "u3": 4838412258.764764, "u4": 4951300212.210368,
    # u5..u7 are absent from the docs example; ibmq_manila's u-channel LO is the target qubit's drive LO
    "u5": 5036989248.286842, "u6": 5066350584.469812, "u7": 4951300212.210368,
}


def _manila_dict() -> dict:
    """The 5-qubit ibmq_manila Hamiltonian dictionary (transmon_model.py:54-115)."""
    h_str = [
        "_SUM[i,0,4,wq{i}/2*(I{i}-Z{i})]",
        "_SUM[i,0,4,delta{i}/2*O{i}*O{i}]",
        "_SUM[i,0,4,-delta{i}/2*O{i}]",
        "_SUM[i,0,4,omegad{i}*X{i}||D{i}]",
    ]
    for (a, b) in _J:
        h_str += [f"jq{a}q{b}*Sp{a}*Sm{b}", f"jq{a}q{b}*Sm{a}*Sp{b}"]
    # cross-drive (control channel) terms, transmon_model.py:82-89
    for (amp_q, op_q, u) in [(1, 0, 0), (0, 1, 1), (2, 1, 2), (1, 2, 3), (3, 2, 4), (4, 3, 6), (2, 3, 5), (3, 4, 7)]:
        h_str.append(f"omegad{amp_q}*X{op_q}||U{u}")
    h_vars = {}
    for q in range(5):
        h_vars[f"wq{q}"] = _WQ[q]
        h_vars[f"delta{q}"] = _DELTA[q]
        h_vars[f"omegad{q}"] = _OMEGAD[q]
    for (a, b), v in _J.items():
        h_vars[f"jq{a}q{b}"] = v
    return {"h_str": h_str, "osc": {}, "qub": {str(q): 3 for q in range(5)}, "vars": h_vars}


MANILA = _manila_dict()


def transmon_hamiltonian_dict(qubit_map: dict, coupling_map: dict) -> dict:
    """String spec built the way TransmonModel.__init__ does (transmon_model.py:139-161).

    qubit_map: {q: (frequency, anharmonicity, drive)}; coupling_map: {(q0, q1): J}.
    """
    q_max = len(qubit_map) - 1
    h_str = [
        f"_SUM[i,0,{q_max},wq{{i}}/2*(I{{i}}-Z{{i}})]",
        f"_SUM[i,0,{q_max},delta{{i}}/2*O{{i}}*O{{i}}]",
        f"_SUM[i,0,{q_max},-delta{{i}}/2*O{{i}}]",
        f"_SUM[i,0,{q_max},omegad{{i}}*X{{i}}||D{{i}}]",
    ]
    h_qub, h_vars = {}, {}
    for q, (freq, anh, drive) in qubit_map.items():
        h_qub[str(q)] = 3
        h_vars[f"wq{q}"] = freq
        h_vars[f"delta{q}"] = anh
        h_vars[f"omegad{q}"] = drive
    for (q0, q1), j in coupling_map.items():
        h_vars[f"jq{q0}q{q1}"] = j
        h_str += [
            f"jq{q0}q{q1}*Sp{q0}*Sm{q1}",
            f"jq{q0}q{q1}*Sp{q1}*Sm{q0}",
            f"omegad{q1}*X{q0}||U{q0}",
            f"omegad{q0}*X{q1}||U{q1}",
        ]
    return {"h_str": h_str, "qub": h_qub, "vars": h_vars}


# --------------------------------------------------------------------------------------
# parse_backend_hamiltonian_dict restatement
# --------------------------------------------------------------------------------------
class _Op:
    """Operator algebra value: '*' between two operators is the matrix product."""

    __array_priority__ = 1000

    def __init__(self, m):
        self.m = np.asarray(m, dtype=complex)

    @staticmethod
    def _mat(x):
        return x.m if isinstance(x, _Op) else x

    def __add__(self, o):
        return _Op(self.m + _Op._mat(o))

    __radd__ = __add__

    def __sub__(self, o):
        return _Op(self.m - _Op._mat(o))

    def __rsub__(self, o):
        return _Op(_Op._mat(o) - self.m)

    def __neg__(self):
        return _Op(-self.m)

    def __mul__(self, o):
        return _Op(self.m @ o.m) if isinstance(o, _Op) else _Op(self.m * o)

    def __rmul__(self, o):
        return _Op(o * self.m)

    def __truediv__(self, o):
        return _Op(self.m / o)


def _local_ops(dim: int) -> dict:
    a = np.diag(np.sqrt(np.arange(1, dim)), k=1).astype(complex)
    ad = a.conj().T
    n = ad @ a
    eye = np.eye(dim, dtype=complex)
    return {"Sm": a, "A": a, "Sp": ad, "C": ad, "O": n, "N": n, "X": a + ad, "Y": -1j * (a - ad),
            "Z": eye - 2 * n, "I": eye}


def _expand_sums(terms: list) -> list:
    out = []
    pat = re.compile(r"_SUM\[(\w+),(-?\d+),(-?\d+),(.*)\]$")
    for term in terms:
        m = pat.match(term.strip())
        if m is None:
            out.append(term.strip())
            continue
        var, lo, hi, body = m.group(1), int(m.group(2)), int(m.group(3)), m.group(4)
        expanded = [body.replace("{" + var + "}", str(i)) for i in range(lo, hi + 1)]
        out += _expand_sums(expanded)
    return out


def parse_hamiltonian_dict(h_dict: dict, subsystem_list=None):
    """-> (static [d,d], operators [K,d,d], channels [K] (sorted, lower-case), dims {label: dim}).

    Subsystem 0 is the least-significant Kronecker factor (qiskit little-endian); only terms
    acting entirely inside ``subsystem_list`` survive; same-channel terms are summed.
    """
    qub = {int(k): int(v) for k, v in h_dict["qub"].items()}
    if subsystem_list is None:
        subsystem_list = sorted(qub)
    subsystem_list = sorted(int(q) for q in subsystem_list)
    dims = {q: qub[q] for q in subsystem_list}
    local = {q: _local_ops(dims[q]) for q in subsystem_list}

    def embed(q, m):
        out = np.array([[1.0 + 0j]])
        for s in subsystem_list:  # ascending label; later labels go to the LEFT
            out = np.kron(m if s == q else np.eye(dims[s]), out)
        return out

    tok = re.compile(r"\b(Sm|Sp|A|C|O|N|X|Y|Z|I)(\d+)\b")
    static = None
    per_channel = {}
    total_dim = int(np.prod([dims[q] for q in subsystem_list]))
    for term in _expand_sums(list(h_dict["h_str"])):
        if "||" in term:
            expr, chan = term.split("||")
            chan = chan.strip().lower()
        else:
            expr, chan = term, None
        used = [(m.group(1), int(m.group(2))) for m in tok.finditer(expr)]
        if not used or any(q not in dims for _, q in used):
            continue
        ns = {f"{name}{q}": _Op(embed(q, local[q][name])) for name, q in used}
        ns.update({k: float(v) for k, v in h_dict["vars"].items()})
        val = eval(expr, {"__builtins__": {}}, ns)  # noqa: S307 - arithmetic on test fixtures only
        mat = val.m if isinstance(val, _Op) else val * np.eye(total_dim)
        if chan is None:
            static = mat if static is None else static + mat
        else:
            per_channel[chan] = per_channel.get(chan, 0) + mat
    if static is None:
        static = np.zeros((total_dim, total_dim), dtype=complex)
    channels = sorted(per_channel)
    operators = np.array([per_channel[c] for c in channels], dtype=complex).reshape(len(channels), total_dim, total_dim)
    return static, operators, channels, dims


def _fix_phase_columns(vecs: np.ndarray) -> np.ndarray:
    """Phase convention shared with the product: largest-|.| component of each column real, positive."""
    vecs = np.array(vecs, dtype=complex)
    for c in range(vecs.shape[1]):
        k = int(np.argmax(np.abs(vecs[:, c])))
        vecs[:, c] *= np.exp(-1j * np.angle(vecs[:, c][k]))
    return vecs


def dressed_decomposition(static_lab: np.ndarray):
    """Dressed eigenbasis ordered by maximum overlap with bare states (Appendix A.4/A.6).

    qiskit-dynamics `_get_dressed_state_decomposition`: eigh, then column c goes to position
    argmax|evec|.  Upstream keeps LAPACK's arbitrary phase; we fix it (see _fix_phase_columns).
    """
    evals, evecs = np.linalg.eigh(static_lab)
    d = len(evals)
    d_evals = np.zeros(d)
    d_states = np.zeros((d, d), dtype=complex)
    seen = set()
    for ev, vec in zip(evals, evecs.T):
        pos = int(np.argmax(np.abs(vec)))
        if pos in seen:
            raise ValueError("dressed-state decomposition: ambiguous overlap")
        seen.add(pos)
        d_states[:, pos] = vec
        d_evals[pos] = ev
    return d_evals, _fix_phase_columns(d_states)


@dataclass
class OracleModel:
    """H(t) = H0 + sum_k s_k(t) H_k in a rotating frame, optional RWA (Appendix A.4).

    ``rotating_frame``: None (lab), Hermitian [d,d] matrix H_f, or real 1-D diagonal of H_f.
    Generator in the frame basis: [G_F]_ij = -i e^{i(e_i-e_j)t} [U^+(H(t)-H_f)U]_ij.
    """

    static: np.ndarray
    operators: np.ndarray
    channels: list
    carrier_freqs: dict
    dt: float
    rotating_frame: np.ndarray | None = None
    rwa_cutoff_freq: float | None = None
    dissipators: np.ndarray | None = None  # [J,d,d] lab-frame static Lindblad operators
    # derived
    frame_evals: np.ndarray = field(init=False)
    frame_basis: np.ndarray = field(init=False)

    def __post_init__(self):
        self.static = np.asarray(self.static, dtype=complex)
        d = self.static.shape[0]
        self.operators = np.asarray(self.operators, dtype=complex).reshape(-1, d, d)
        self.dim = d
        rf = self.rotating_frame
        if rf is None:
            self.frame_evals = np.zeros(d)
            self.frame_basis = np.eye(d, dtype=complex)
            h_f = np.zeros((d, d), dtype=complex)
        else:
            rf = np.asarray(rf)
            if rf.ndim == 1:
                self.frame_evals = rf.real.astype(float)
                self.frame_basis = np.eye(d, dtype=complex)
                h_f = np.diag(self.frame_evals).astype(complex)
            else:
                ev, U = np.linalg.eigh(rf)
                self.frame_evals, self.frame_basis = ev, U.astype(complex)
                h_f = rf.astype(complex)
        U = self.frame_basis
        self.static_fb = U.conj().T @ (self.static - h_f) @ U
        self.ops_fb = np.array([U.conj().T @ o @ U for o in self.operators]).reshape(-1, d, d)
        self.nu = np.array([self.carrier_freqs[c] for c in self.channels], dtype=float)
        self.bohr = self.frame_evals[:, None] - self.frame_evals[None, :]  # e_i - e_j  [rad/s]
        # RWA masks (element kept iff |+-nu + bohr/2pi| < cutoff); pos part multiplies z, neg part z*
        if self.rwa_cutoff_freq is None:
            self.pos = 0.5 * self.ops_fb
            self.neg = 0.5 * self.ops_fb
            self.static_eff = self.static_fb
        else:
            c = float(self.rwa_cutoff_freq)
            f = self.bohr / (2 * np.pi)
            self.static_eff = np.where(np.abs(f) < c, self.static_fb, 0)
            self.pos = np.array([0.5 * np.where(np.abs(nu + f) < c, o, 0) for nu, o in zip(self.nu, self.ops_fb)]).reshape(-1, d, d)
            self.neg = np.array([0.5 * np.where(np.abs(-nu + f) < c, o, 0) for nu, o in zip(self.nu, self.ops_fb)]).reshape(-1, d, d)
        if self.dissipators is not None:
            self.dissipators = np.asarray(self.dissipators, dtype=complex).reshape(-1, d, d)
            self.diss_fb = np.array([U.conj().T @ L @ U for L in self.dissipators]).reshape(-1, d, d)

    # lab-frame static Hamiltonian (what DynamicsBackend diagonalises for the dressed basis)
    def dressed(self):
        return dressed_decomposition(self.static)

    def hamiltonian_fb(self, t, z):
        """Frame-basis H_F(t) for complex channel values z = env(t) e^{i 2 pi nu t}; z: [..., K].

        Returns [..., d, d].  Without RWA this equals S + sum_k Re(z_k) G_k (real signals).
        """
        z = np.asarray(z, dtype=complex)
        h = self.static_eff + np.einsum("...k,kij->...ij", z, self.pos) + np.einsum("...k,kij->...ij", z.conj(), self.neg)
        return np.exp(1j * self.bohr * t) * h

    def generator_fb(self, t, z):
        return -1j * self.hamiltonian_fb(t, z)
