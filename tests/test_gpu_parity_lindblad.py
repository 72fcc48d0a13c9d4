"""GPU parity for the structured Lindblad RK4 path (BASELINE cfg 4) against the oracle's dense restatement."""
import os

import numpy as np
import pytest
import torch

import oracle as O
from casq_b200 import engine
from casq_b200._native import NativeModel
from casq_b200.common.exceptions import CasqError

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DT = O.MANILA_DT
TOL_RHO = 1e-9


def _t1t2_ops(n_qubits, T1=100e-6, T2=80e-6, scale=1.0):
    """L1 = sqrt(1/T1) a, Lphi = sqrt(2 gamma_phi) a^+ a per qubit (SURVEY §8(d) cfg 4); `scale` shortens the times."""
    a = np.diag(np.sqrt(np.arange(1, 3)), k=1).astype(complex)
    n = a.conj().T @ a
    g1, gphi = scale / T1, scale * (1 / T2 - 1 / (2 * T1))
    out = []
    for q in range(n_qubits):
        def emb(m_, q=q):
            mats = [np.eye(3)] * n_qubits
            mats[q] = m_
            full = np.array([[1.0 + 0j]])
            for s in range(n_qubits):  # subsystem 0 right-most
                full = np.kron(mats[s], full)
            return full
        out += [np.sqrt(g1) * emb(a), np.sqrt(2 * gphi) * emb(n)]
    return np.array(out)


def _two_transmon():
    v = O.MANILA["vars"]
    hd = O.transmon_hamiltonian_dict({0: (v["wq0"], v["delta0"], v["omegad0"]), 1: (v["wq1"], v["delta1"], v["omegad1"])},
                                     {(0, 1): v["jq0q1"]})
    return O.parse_hamiltonian_dict(hd, None)


def test_lindblad_d9_golden():
    g = np.load(os.path.join(G, "lindblad_d9.npz"))
    st, ops, ch, _ = _two_transmon()
    Ls = _t1t2_ops(2, float(g["T1"]), float(g["T2"]))
    nm = NativeModel(st, ops, [O.MANILA_CARRIERS[c] for c in ch], DT, rotating_frame=np.diag(st).real, rwa_cutoff_freq=3.0e9)
    nm.set_dissipators(Ls)
    p = dict(amp=g["amp"], angle=0.0, sigma=float(g["sigma"]), beta=float(g["beta"]))
    dur = int(g["duration"])
    rho0 = np.zeros(9, dtype=complex)
    rho0[0] = 1
    rho, pops = engine.solve_lindblad_rk4(nm, [("drag", ch.index("d0"), 0, dur)], engine.pack_params([p], device="cuda"), rho0,
                                          (0.0, dur * DT), DT / 4)
    assert np.abs(rho[:, -1].cpu().numpy() - g["rho"]).max() < TOL_RHO
    pp = pops[:, -1].cpu().numpy()
    assert np.abs(pp / pp.sum(-1, keepdims=True) - g["pops"]).max() < TOL_RHO


@pytest.mark.parametrize("rwa,frame_shift,strong", [(None, 0.0, False), (3.0e9, 0.0, True), (None, 2e8, True)])
def test_lindblad_d3_matches_oracle(rwa, frame_shift, strong):
    """Single transmon, T1/T2 (optionally 1e3 x stronger so that dissipation matters), t_eval, batched rho0."""
    rng = np.random.default_rng(8)
    st, ops, ch, _ = O.parse_hamiltonian_dict(O.MANILA, [0])
    Ls = _t1t2_ops(1, scale=1e3 if strong else 1.0)
    rf = np.diag(st).real + frame_shift
    om = O.OracleModel(st, ops, ch, O.MANILA_CARRIERS, DT, rotating_frame=rf, rwa_cutoff_freq=rwa, dissipators=Ls)
    nm = NativeModel(st, ops, [O.MANILA_CARRIERS[c] for c in ch], DT, rotating_frame=rf, rwa_cutoff_freq=rwa)
    nm.set_dissipators(Ls)
    B, dur = 5, 20
    p = dict(amp=rng.uniform(0.2, 0.9, B), angle=rng.uniform(-1, 1, B), sigma=5.0, beta=rng.uniform(-2, 2, B))
    samples = O.channel_samples(ch, [O.PulseSpec("drag", "d0", dur)], [p])
    psi = rng.normal(size=(B, 3)) + 1j * rng.normal(size=(B, 3))
    psi /= np.linalg.norm(psi, axis=1, keepdims=True)
    rho0 = np.einsum("bi,bj->bij", psi, psi.conj())
    te = np.linspace(0, dur * DT, 4)
    sub = 40 if rwa is None else 8
    want = np.stack([O.lindblad_rk4_solve(om, samples[b:b + 1], rho0[b], (0.0, dur * DT), DT / sub, te)[1][0] for b in range(B)])
    rho, pops = engine.solve_lindblad_rk4(nm, [("drag", 0, 0, dur)], engine.pack_params([p], device="cuda"), rho0, (0.0, dur * DT),
                                          DT / sub, te)
    got = rho.cpu().numpy()
    assert got.shape == (B, 4, 3, 3) and np.abs(got - want).max() < TOL_RHO
    tr = np.trace(got, axis1=-2, axis2=-1)
    assert np.abs(tr - 1).max() < 1e-9 and np.abs(got - np.conj(np.swapaxes(got, -1, -2))).max() < 1e-12
    ref_pops = np.real(np.einsum("btii->bti", O.postprocess_density_matrices(om, te, want, normalize=False)))
    assert np.abs(pops.cpu().numpy() - ref_pops).max() < TOL_RHO


def test_lindblad_amplitude_damping_closed_form():
    """No drive: P1(t) = exp(-t/T1) for |1><1|, coherence decays at 1/T2 (gamma_1/2 + gamma_phi)."""
    st, ops, ch, _ = O.parse_hamiltonian_dict(O.MANILA, [0])
    T1, T2 = 40e-9, 30e-9
    nm = NativeModel(st, ops, [O.MANILA_CARRIERS[c] for c in ch], DT, rotating_frame=np.diag(st).real)
    nm.set_dissipators(_t1t2_ops(1, T1, T2))
    psi = np.array([1, 1, 0]) / np.sqrt(2)
    T = 64 * DT
    rho, pops = engine.solve_lindblad_rk4(nm, [("constant", 0, 0, 64)], engine.pack_params([dict(amp=0.0)], device="cuda"), psi,
                                          (0.0, T), DT / 2)
    r = rho[0, -1].cpu().numpy()
    assert abs(r[1, 1].real - 0.5 * np.exp(-T / T1)) < 1e-9
    assert abs(abs(r[0, 1]) - 0.5 * np.exp(-T / T2)) < 1e-9
    assert abs(np.trace(r) - 1) < 1e-12


def test_lindblad_needs_diagonal_frame():
    st, ops, ch, _ = _two_transmon()
    nm = NativeModel(st, ops, [O.MANILA_CARRIERS[c] for c in ch], DT, rotating_frame=st)  # full-matrix frame: dressed basis
    nm.set_dissipators(_t1t2_ops(2))
    with pytest.raises(CasqError, match="diagonal rotating frame"):
        engine.solve_lindblad_rk4(nm, [("drag", 0, 0, 8)], engine.pack_params([dict(amp=0.1, sigma=2.0, beta=0.0)], device="cuda"),
                                  np.eye(9)[0], (0.0, 8 * DT), DT)


def test_lindblad_d27_three_transmons_vs_oracle():
    st, ops, ch, _ = O.parse_hamiltonian_dict(O.MANILA, [0, 1, 2])
    Ls = _t1t2_ops(3, scale=1e3)
    rf = np.diag(st).real
    om = O.OracleModel(st, ops, ch, O.MANILA_CARRIERS, DT, rotating_frame=rf, rwa_cutoff_freq=3.0e9, dissipators=Ls)
    nm = NativeModel(st, ops, [O.MANILA_CARRIERS[c] for c in ch], DT, rotating_frame=rf, rwa_cutoff_freq=3.0e9)
    nm.set_dissipators(Ls)
    p = dict(amp=np.array([0.3, 0.8]), angle=0.0, sigma=4.0, beta=1.0)
    samples = O.channel_samples(ch, [O.PulseSpec("drag", "d1", 12)], [p])
    rho0 = np.zeros(27, dtype=complex)
    rho0[0] = 1
    _, want = O.lindblad_rk4_solve(om, samples, rho0, (0.0, 12 * DT), DT / 4)
    rho, _ = engine.solve_lindblad_rk4(nm, [("drag", ch.index("d1"), 0, 12)], engine.pack_params([p], device="cuda"), rho0,
                                       (0.0, 12 * DT), DT / 4, want_pops=False)
    assert np.abs(rho.cpu().numpy() - want).max() < TOL_RHO


def test_lindblad_parts_on_several_streams_are_bit_identical(monkeypatch):
    """Large batches are cut into up to four parts that run on their own streams (their grids fill each other's tails).
    Forcing 1, 3 and 4 parts on a ragged batch of 7, with batched rho0, t_eval and both outputs, must give identical bits."""
    rng = np.random.default_rng(31)
    st, ops, ch, _ = _two_transmon()
    Ls = _t1t2_ops(2, scale=1e3)
    nm = NativeModel(st, ops, [O.MANILA_CARRIERS[c] for c in ch], DT, rotating_frame=np.diag(st).real, rwa_cutoff_freq=3.0e9)
    nm.set_dissipators(Ls)
    B, dur = 7, 10
    p = dict(amp=rng.uniform(0.2, 0.9, B), angle=rng.uniform(-1, 1, B), sigma=3.0, beta=rng.uniform(-2, 2, B))
    psi = rng.normal(size=(B, 9)) + 1j * rng.normal(size=(B, 9))
    psi /= np.linalg.norm(psi, axis=1, keepdims=True)
    rho0 = np.einsum("bi,bj->bij", psi, psi.conj())
    te = np.linspace(0, dur * DT, 3)
    out = {}
    for parts in ("1", "3", "4"):
        monkeypatch.setenv("CASQ_LB_PARTS", parts)
        rho, pops = engine.solve_lindblad_rk4(nm, [("drag", ch.index("d0"), 0, dur)], engine.pack_params([p], device="cuda"), rho0,
                                              (0.0, dur * DT), DT / 4, t_eval=te)
        torch.cuda.synchronize()
        out[parts] = (rho.clone(), pops.clone())
    for parts in ("3", "4"):
        assert torch.equal(out["1"][0], out[parts][0]) and torch.equal(out["1"][1], out[parts][1])


def _chain_problem(qubits, t1t2_scale, rwa=True):
    st, ops, ch, _ = O.parse_hamiltonian_dict(O.MANILA, list(qubits))
    Ls = _t1t2_ops(len(qubits), scale=t1t2_scale)
    rf = np.diag(st).real
    cut = 0.75 * O.MANILA_CARRIERS["d0"] if rwa else None
    om = O.OracleModel(st, ops, ch, O.MANILA_CARRIERS, DT, rotating_frame=rf, rwa_cutoff_freq=cut, dissipators=Ls)
    nm = NativeModel(st, ops, [O.MANILA_CARRIERS[c] for c in ch], DT, rotating_frame=rf, rwa_cutoff_freq=cut)
    nm.set_dissipators(Ls)
    return om, nm, ch


def _random_density_matrices(rng, B, d, rank=3):
    x = rng.normal(size=(B, d, rank)) + 1j * rng.normal(size=(B, d, rank))
    rho = x @ np.conj(np.swapaxes(x, -1, -2))
    return rho / np.trace(rho, axis1=-2, axis2=-1).real[:, None, None]


@pytest.mark.parametrize("parts", ["1", "4"])
@pytest.mark.parametrize("rho0_kind", ["ground", "random"])
def test_lindblad_d81_off_diagonal_tile_pairs_vs_oracle(parts, rho0_kind, monkeypatch):
    """Four transmons (d = 81): more than one tile per side, so the off-diagonal tile-pair path (left gathers of the
    mirrored tile, transposed read, mirror store) is compared ELEMENT BY ELEMENT with the dense oracle -- 16 RK4 steps,
    ragged batch, one and four stream parts, a random mixed rho0 so that every tile is non-zero from step 0, T1/T2 1e3 x
    stronger so that the two-sided terms matter at 1e-9."""
    monkeypatch.setenv("CASQ_LB_PARTS", parts)
    rng = np.random.default_rng(81)
    om, nm, ch = _chain_problem([0, 1, 2, 3], 1e3)
    B, dur = 5, 8
    p = dict(amp=rng.uniform(0.2, 0.9, B), angle=rng.uniform(-1, 1, B), sigma=3.0, beta=rng.uniform(-2, 2, B))
    samples = O.channel_samples(ch, [O.PulseSpec("drag", "d0", dur)], [p])
    span = (0.0, 4 * DT)
    te = np.array([0.0, 1.5, 4.0]) * DT
    if rho0_kind == "ground":
        rho0 = np.zeros((81, 81), dtype=complex)
        rho0[0, 0] = 1
        want = O.lindblad_rk4_solve(om, samples, rho0, span, DT / 4, te)[1]
    else:
        rho0 = _random_density_matrices(rng, B, 81)
        want = np.concatenate([O.lindblad_rk4_solve(om, samples[b:b + 1], rho0[b], span, DT / 4, te)[1] for b in range(B)])
    rho, pops = engine.solve_lindblad_rk4(nm, [("drag", ch.index("d0"), 0, dur)], engine.pack_params([p], device="cuda"), rho0,
                                          span, DT / 4, te)
    got = rho.cpu().numpy()
    assert got.shape == (B, 3, 81, 81)
    assert np.abs(got - want).max() < TOL_RHO
    ref_pops = np.real(np.einsum("btii->bti", O.postprocess_density_matrices(om, te, want, normalize=False)))
    assert np.abs(pops.cpu().numpy() - ref_pops).max() < TOL_RHO


@pytest.mark.parametrize("scale,rho0_kind", [(1.0, "ground"), (1e3, "random")])
def test_lindblad_d243_cfg4_operators_vs_oracle(scale, rho0_kind, monkeypatch):
    """BASELINE cfg 4's own operators (full ibmq_manila, d = 243, RWA cutoff 0.75 d0, T1 = 100 us / T2 = 80 us per qubit,
    DRAG on d0) against the dense oracle, element by element, for 20 RK4 steps: every one of the tile pairs of the
    production kernel is exercised; the second case has a random mixed rho0 and 1e3 x faster T1/T2."""
    rng = np.random.default_rng(243)
    om, nm, ch = _chain_problem([0, 1, 2, 3, 4], scale)
    B, dur = 2, 8
    p = dict(amp=np.array([0.3, 0.85]), angle=np.array([0.0, 0.7]), sigma=3.0, beta=np.array([1.0, -2.0]))
    samples = O.channel_samples(ch, [O.PulseSpec("drag", "d0", dur)], [p])
    span = (0.0, 5 * DT)
    if rho0_kind == "ground":
        rho0 = np.zeros((243, 243), dtype=complex)
        rho0[0, 0] = 1
        want = O.lindblad_rk4_solve(om, samples, rho0, span, DT / 4)[1]
    else:
        rho0 = _random_density_matrices(rng, B, 243)
        want = np.concatenate([O.lindblad_rk4_solve(om, samples[b:b + 1], rho0[b], span, DT / 4)[1] for b in range(B)])
    monkeypatch.setenv("CASQ_LB_CHECK_PLAN", "1")  # the library verifies that every gather piece stays inside the work buffers
    for parts in ("1", "2"):
        monkeypatch.setenv("CASQ_LB_PARTS", parts)
        rho, pops = engine.solve_lindblad_rk4(nm, [("drag", ch.index("d0"), 0, dur)], engine.pack_params([p], device="cuda"),
                                              rho0, span, DT / 4)
        got = rho.cpu().numpy()
        assert np.abs(got - want).max() < TOL_RHO, parts
        assert np.abs(got[:, -1] - np.conj(np.swapaxes(got[:, -1], -1, -2))).max() < 1e-13
        ref_pops = np.real(np.einsum("btii->bti", O.postprocess_density_matrices(om, np.array(span), want, normalize=False)))
        assert np.abs(pops.cpu().numpy() - ref_pops).max() < TOL_RHO


@pytest.mark.parametrize("variant", ["CASQ_LB_NO_TMA", "CASQ_LB_GENERIC"])
def test_lindblad_other_kernel_variants_keep_parity(variant, monkeypatch):
    """The default chain kernel stages its operands with bulk asynchronous copies (TMA).  The same stage with plain loads
    (CASQ_LB_NO_TMA) and the generic shift-class kernel (CASQ_LB_GENERIC: what a model that is not a chain of three-level
    sites runs) must give the same answers: d = 81 against the dense oracle, and against the default path to rounding."""
    rng = np.random.default_rng(811)
    om, nm, ch = _chain_problem([0, 1, 2, 3], 1e3)
    B, dur = 3, 8
    p = dict(amp=rng.uniform(0.2, 0.9, B), angle=rng.uniform(-1, 1, B), sigma=3.0, beta=rng.uniform(-2, 2, B))
    samples = O.channel_samples(ch, [O.PulseSpec("drag", "d0", dur)], [p])
    rho0 = _random_density_matrices(rng, B, 81)
    span = (0.0, 3 * DT)
    want = np.concatenate([O.lindblad_rk4_solve(om, samples[b:b + 1], rho0[b], span, DT / 4)[1] for b in range(B)])
    args = (nm, [("drag", ch.index("d0"), 0, dur)], engine.pack_params([p], device="cuda"), rho0, span, DT / 4)
    default, _ = engine.solve_lindblad_rk4(*args)
    monkeypatch.setenv(variant, "1")
    other, _ = engine.solve_lindblad_rk4(*args)
    assert np.abs(other.cpu().numpy() - want).max() < TOL_RHO
    assert float((other - default).abs().max()) < 1e-13


@pytest.mark.parametrize("drive,rwa", [("d3", True), ("d2", True), ("u2", True), ("d0", False)])
def test_lindblad_chain_kernel_other_drives_and_no_rwa(drive, rwa):
    """The chain kernel's remaining branches against the dense oracle at d = 81: a drive on site 3 (a tile-level ladder term),
    on site 2 (the digit the three warps of a block split), on a control channel (X of one site at another site's
    frequency), and the lab-frame-carrier case without RWA (real signal, both rotating terms)."""
    rng = np.random.default_rng(7)
    om, nm, ch = _chain_problem([0, 1, 2, 3], 1e3, rwa=rwa)
    B, dur = 3, 8
    p = dict(amp=rng.uniform(0.3, 0.9, B), angle=rng.uniform(-1, 1, B), sigma=3.0, beta=rng.uniform(-2, 2, B))
    samples = O.channel_samples(ch, [O.PulseSpec("drag", drive, dur)], [p])
    rho0 = _random_density_matrices(rng, B, 81)
    span = (0.0, 2 * DT)
    sub = 4 if rwa else 40
    want = np.concatenate([O.lindblad_rk4_solve(om, samples[b:b + 1], rho0[b], span, DT / sub)[1] for b in range(B)])
    rho, pops = engine.solve_lindblad_rk4(nm, [("drag", ch.index(drive), 0, dur)], engine.pack_params([p], device="cuda"), rho0,
                                          span, DT / sub)
    assert np.abs(rho.cpu().numpy() - want).max() < TOL_RHO
    ref_pops = np.real(np.einsum("btii->bti", O.postprocess_density_matrices(om, np.array(span), want, normalize=False)))
    assert np.abs(pops.cpu().numpy() - ref_pops).max() < TOL_RHO


def test_lindblad_chain_kernel_without_dissipators_equals_state_vector_solve():
    """No noise model at all (zero dissipators): rho(t) = |psi(t)><psi(t)| of the state-vector RK4 to the integrators' accuracy."""
    st, ops, ch, _ = O.parse_hamiltonian_dict(O.MANILA, [0, 1, 2])
    rf = np.diag(st).real
    cut = 0.75 * O.MANILA_CARRIERS["d0"]
    nm = NativeModel(st, ops, [O.MANILA_CARRIERS[c] for c in ch], DT, rotating_frame=rf, rwa_cutoff_freq=cut)
    nm.set_dissipators(np.zeros((0, 27, 27), dtype=complex))
    p = dict(amp=np.array([0.4, 0.9]), angle=0.0, sigma=6.0, beta=1.0)
    params = engine.pack_params([p], device="cuda")
    slots = [("drag", ch.index("d0"), 0, 24)]
    psi0 = np.zeros(27, dtype=complex)
    psi0[0] = 1
    span = (0.0, 24 * DT)
    rho, _ = engine.solve_lindblad_rk4(nm, slots, params, psi0, span, DT / 8)
    y = engine.solve_sv_rk4(nm, slots, params, psi0, span, DT / 8)[:, -1]
    outer = torch.einsum("bi,bj->bij", y, y.conj())
    assert float((rho[:, -1] - outer).abs().max()) < 5e-6  # two different RK4 discretisations of the same flow (measured 4.9e-7)
    assert float((torch.diagonal(rho[:, -1], dim1=-2, dim2=-1).sum(-1).real - 1).abs().max()) < 1e-9
