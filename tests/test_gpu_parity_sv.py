"""GPU parity tests proper: the CUDA path, called through the C ABI, against the oracle on the same seeded
inputs (bit-level agreement is not expected for fp64 transcendental code; the bar is north_star's 1e-6 on
state vectors -- these tests hold the kernels to 1e-9 or better where both sides run the same scheme)."""
import os

import numpy as np
import pytest
import torch

import oracle as O
from casq_b200 import engine
from casq_b200._native import NativeModel
from casq_b200.backends import BackendLibrary, PulseBackend, build
from casq_b200.common.exceptions import CasqError
from casq_b200.gates import DragPulseGate, GaussianPulseGate, GaussianSquarePulseGate, PulseCircuit
from casq_b200.models import ControlModel, TransmonModel

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DT = O.MANILA_DT
TOL_STATE = 1e-9  # same RK4 scheme on both sides; north_star's bar is 1e-6


def _models(subsystems=(0,), frame="full", rwa=None, h_dict=None, carriers=None):
    h_dict = O.MANILA if h_dict is None else h_dict
    carriers = O.MANILA_CARRIERS if carriers is None else carriers
    st, ops, ch, _ = O.parse_hamiltonian_dict(h_dict, list(subsystems) if subsystems is not None else None)
    rf = {"full": st, "diag": np.diag(st).real, "none": None, "shifted": np.diag(st).real + 2e8}[frame]
    om = O.OracleModel(st, ops, ch, carriers, DT, rotating_frame=rf, rwa_cutoff_freq=rwa)
    nm = NativeModel(st, ops, [carriers[c] for c in ch], DT, rotating_frame=rf, rwa_cutoff_freq=rwa)
    return om, nm, ch


def _rand_params(rng, kind, B, duration):
    p = dict(amp=rng.uniform(0.05, 0.9, B), angle=rng.uniform(-np.pi, np.pi, B))
    if kind != "constant":
        p["sigma"] = rng.uniform(duration / 10, duration / 2, B)
    if "square" in kind:
        p["width"] = rng.uniform(0, duration * 0.7, B)
    if "drag" in kind:
        p["beta"] = rng.uniform(-3, 3, B)
    return p


@pytest.mark.parametrize("kind", ["constant", "gaussian", "drag", "gaussian_square", "gaussian_square_drag"])
def test_envelope_sampler_matches_oracle(kind):
    rng = np.random.default_rng(11)
    B, dur = 257, 48
    p = _rand_params(rng, kind, B, dur)
    want = O.envelope_samples(kind, dur, **p)
    got = engine.envelope_sample(kind, dur, engine.pack_params([p], device="cuda")[0]).cpu().numpy()
    assert np.abs(got - want).max() < 1e-14


def test_envelope_sampler_golden_and_edge_cases():
    g = np.load(os.path.join(G, "envelopes.npz"))
    for kind, keys in [("constant", ()), ("gaussian", ("sigma",)), ("drag", ("sigma", "beta")),
                       ("gaussian_square", ("sigma", "width")), ("gaussian_square_drag", ("sigma", "width", "beta"))]:
        p = dict(amp=g["p_amp"], angle=g["p_angle"], **{k: g[f"p_{k}"] for k in keys})
        got = engine.envelope_sample(kind, int(g["duration"]), engine.pack_params([p], device="cuda")[0]).cpu().numpy()
        assert np.abs(got - g[kind]).max() < 1e-14
    # width = 0 and width = duration, duration = 1, empty batch
    for w in (0.0, 16.0):
        p = dict(amp=0.5, angle=0.0, sigma=3.0, width=w)
        got = engine.envelope_sample("gaussian_square", 16, engine.pack_params([p], device="cuda")[0]).cpu().numpy()
        assert np.abs(got - O.envelope_samples("gaussian_square", 16, **p)).max() < 1e-15
    one = engine.envelope_sample("gaussian", 1, engine.pack_params([dict(amp=0.3, sigma=2.0)], device="cuda")[0])
    assert np.abs(one.cpu().numpy() - O.envelope_samples("gaussian", 1, amp=0.3, sigma=2.0)).max() < 1e-15
    empty = engine.envelope_sample("gaussian", 4, torch.zeros((6, 0), dtype=torch.float64, device="cuda"))
    assert empty.shape == (0, 4)


@pytest.mark.parametrize("kind", ["constant", "gaussian", "drag", "gaussian_square", "gaussian_square_drag"])
@pytest.mark.parametrize("frame,rwa", [("full", None), ("none", None), ("shifted", None), ("full", 3.0e9), ("shifted", 3.0e9)])
def test_rk4_d3_matches_oracle(kind, frame, rwa):
    rng = np.random.default_rng(5)
    om, nm, ch = _models((0,), frame, rwa)
    B, dur = 33, 40
    p = _rand_params(rng, kind, B, dur)
    samples = O.channel_samples(ch, [O.PulseSpec(kind, "d0", dur)], [p])
    span = (0.0, dur * DT)
    te = np.linspace(0, dur * DT, 6)
    sub = 40 if frame != "none" else 160
    _, want = O.rk4_solve(om, samples, np.array([1, 0, 0]), span, DT / sub, te)
    got = engine.solve_sv_rk4(nm, [(kind, ch.index("d0"), 0, dur)], engine.pack_params([p], device="cuda"),
                              np.array([1, 0, 0]), span, DT / sub, te).cpu().numpy()
    assert got.shape == (B, 6, 3)
    assert np.abs(got - want).max() < TOL_STATE


def test_rk4_golden_fixture_and_postprocess():
    g = np.load(os.path.join(G, "rk4_drag_d3.npz"))
    om, nm, ch = _models((0,), "full")
    p = dict(amp=g["amp"], angle=0.0, sigma=float(g["sigma"]), beta=g["beta"])
    raw = engine.solve_sv_rk4(nm, [("drag", 0, 0, int(g["duration"]))], engine.pack_params([p], device="cuda"),
                              np.array([1, 0, 0]), (0.0, int(g["t_final"]) * DT), DT / 40, g["t_eval"])
    assert np.abs(raw.cpu().numpy() - g["raw"]).max() < TOL_STATE
    states, probs, pops = engine.postprocess_sv(nm, g["t_eval"], raw)
    assert np.abs(states.cpu().numpy() - g["states"]).max() < TOL_STATE
    pr = np.abs(g["states"]) ** 2
    assert np.abs(probs.cpu().numpy() - pr).max() < TOL_STATE
    assert np.abs(pops.cpu().numpy()[..., 0] - pr[..., 0]).max() < TOL_STATE
    assert np.abs(pops.cpu().numpy()[..., 1] - pr[..., 1:].sum(-1)).max() < TOL_STATE


def test_rk4_two_level_and_four_level_systems():
    rng = np.random.default_rng(9)
    for dims, d in (({"0": 2}, 2), ({"0": 2, "1": 2}, 4), ({"0": 4}, 4)):
        terms = ["_SUM[i,0,{n},wq{{i}}/2*(I{{i}}-Z{{i}})]".format(n=len(dims) - 1), "_SUM[i,0,{n},om{{i}}*X{{i}}||D{{i}}]".format(n=len(dims) - 1)]
        variables = {"wq0": 2 * np.pi * 4.9e9, "om0": 6e8, "wq1": 2 * np.pi * 5.1e9, "om1": 5e8, "j": 4e7, "al": -2e9}
        if len(dims) == 2:
            terms += ["j*Sp0*Sm1", "j*Sm0*Sp1", "om0*Y1||U0"]
        if d == 4 and len(dims) == 1:
            terms += ["al/2*O0*O0", "-al/2*O0"]
        hd = {"h_str": terms, "qub": dims, "vars": variables}
        carriers = {"d0": 4.9e9, "d1": 5.1e9, "u0": 5.1e9}
        for frame in ("full", "diag"):
            om, nm, ch = _models(None, frame, None, hd, carriers)
            assert om.dim == d
            B, dur = 19, 24
            specs = [O.PulseSpec("drag", "d0", dur)]
            slots = [("drag", ch.index("d0"), 0, dur)]
            ps = [_rand_params(rng, "drag", B, dur)]
            if "u0" in ch:  # second active channel, overlapping in time
                specs.append(O.PulseSpec("gaussian_square", "u0", 16, start=4))
                slots.append(("gaussian_square", ch.index("u0"), 4, 16))
                ps.append(_rand_params(rng, "gaussian_square", B, 16))
            samples = O.channel_samples(ch, specs, ps)
            y0 = rng.normal(size=(B, d)) + 1j * rng.normal(size=(B, d))
            y0 /= np.linalg.norm(y0, axis=1, keepdims=True)
            span = (0.0, dur * DT)
            _, want = O.rk4_solve(om, samples, y0, span, DT / 50)
            got = engine.solve_sv_rk4(nm, slots, engine.pack_params(ps, device="cuda"), y0, span, DT / 50).cpu().numpy()
            assert np.abs(got - want).max() < TOL_STATE, (dims, frame)


def test_rk4_multi_slot_same_channel_and_ragged_grid():
    """Two pulses on one channel (second starts later), max_dt that does not divide dt, t_eval off the grid."""
    rng = np.random.default_rng(21)
    om, nm, ch = _models((0,), "full")
    B = 8
    specs = [O.PulseSpec("gaussian", "d0", 16, start=0), O.PulseSpec("drag", "d0", 20, start=20)]
    ps = [_rand_params(rng, "gaussian", B, 16), _rand_params(rng, "drag", B, 20)]
    samples = O.channel_samples(ch, specs, ps)
    span = (0.0, 40 * DT)
    te = np.array([0.31, 7.77, 20.0, 33.3]) * DT
    _, want = O.rk4_solve(om, samples, np.array([0, 1, 0]), span, DT / 37.3, te)
    got = engine.solve_sv_rk4(nm, [("gaussian", 0, 0, 16), ("drag", 0, 20, 20)], engine.pack_params(ps, device="cuda"),
                              np.array([0, 1, 0]), span, DT / 37.3, te).cpu().numpy()
    assert got.shape == (B, 4, 3) and np.abs(got - want).max() < TOL_STATE


def test_rk4_full_size_properties():
    """BASELINE cfg 2 at full size (1e5 points, S = 10280): size-independent checks."""
    om, nm, ch = _models((0,), "full")
    beta, amp = np.linspace(-5, 5, 250), np.linspace(0.02, 0.5, 400)
    bb, aa = np.meshgrid(beta, amp, indexing="ij")
    p = dict(amp=aa.ravel(), angle=0.0, sigma=64.0, beta=bb.ravel())
    params = engine.pack_params([p], device="cuda")
    span = (0.0, 257 * DT)
    slots = [("drag", 0, 0, 256)]
    y = engine.solve_sv_rk4(nm, slots, params, np.array([1, 0, 0]), span, DT / 40)
    assert y.shape == (100000, 2, 3)
    yf = y[:, 1]
    norms = torch.linalg.vector_norm(yf, dim=1)
    assert float((norms - 1).abs().max()) < 1e-5  # RK4 is not exactly unitary; drift is O(S h^5)
    # linearity in the initial state: solve(a|0> + b|1>) = a solve(|0>) + b solve(|1>)
    sel = torch.arange(0, 100000, 997, device="cuda")
    sub = params[:, :, sel].contiguous()
    y1 = engine.solve_sv_rk4(nm, slots, sub, np.array([0, 1, 0]), span, DT / 40)[:, 1]
    a, b = 0.6, 0.8j
    ymix = engine.solve_sv_rk4(nm, slots, sub, np.array([a, b, 0]), span, DT / 40)[:, 1]
    assert float((ymix - (a * yf[sel] + b * y1)).abs().max()) < 1e-12
    # beta -> -beta with a real envelope conjugates nothing physical: P1 is symmetric only approximately,
    # but the amp = 0.02..0.5 sweep must contain a pi-pulse (P1 > 0.99) near amp ~ 0.111 at beta = 0
    p1 = (yf[:, 1].abs() ** 2).reshape(250, 400)
    row = p1[125, :150]  # beta ~ 0.02, amp < 0.2 (the 3 pi pulse near amp 0.33 also inverts)
    k = int(row.argmax())
    assert float(row[k]) > 0.99 and 0.10 < amp[k] < 0.12
    # oracle spot check at full S on 3 points
    idx = np.array([0, 50321, 99999])
    samples = O.channel_samples(ch, [O.PulseSpec("drag", "d0", 256)], [dict(amp=p["amp"][idx], angle=0.0, sigma=64.0, beta=p["beta"][idx])])
    _, want = O.rk4_solve(om, samples, np.array([1, 0, 0]), span, DT / 40)
    assert np.abs(y.cpu().numpy()[idx] - want).max() < TOL_STATE


def test_dopri5_matches_oracle_and_golden():
    g = np.load(os.path.join(G, "dopri5_drag_d3.npz"))
    om, nm, ch = _models((0,), "full")
    p = dict(amp=float(g["amp"]), angle=0.0, sigma=g["sigma"], beta=g["beta"])
    out, nsteps, status = engine.solve_sv_dopri5(nm, [("drag", 0, 0, int(g["duration"]))], engine.pack_params([p], device="cuda"),
                                                 np.array([1, 0, 0]), (0.0, int(g["duration"]) * DT), rtol=1e-8, atol=1e-6,
                                                 hmax=DT, return_stats=True)
    assert int(status.abs().sum()) == 0
    # same controller => same accept/reject sequence; allow a couple of borderline decisions to differ
    assert np.abs(nsteps.cpu().numpy() - g["steps"]).max() <= 0.02 * g["steps"].max()
    assert np.abs(out.cpu().numpy() - g["raw"]).max() < 1e-6


@pytest.mark.parametrize("frame,rwa", [("full", None), ("shifted", None), ("full", 3.0e9)])
def test_dopri5_dense_output_and_frames(frame, rwa):
    rng = np.random.default_rng(3)
    om, nm, ch = _models((0,), frame, rwa)
    B, dur = 6, 24
    p = _rand_params(rng, "drag", B, dur)
    samples = O.channel_samples(ch, [O.PulseSpec("drag", "d0", dur)], [p])
    span = (0.0, dur * DT)
    te = np.linspace(0, dur * DT, 7)
    _, want, steps = O.dopri5_jax_solve(om, samples, np.array([1, 0, 0]), span, te, rtol=1e-9, atol=1e-9)
    got, ns, status = engine.solve_sv_dopri5(nm, [("drag", 0, 0, dur)], engine.pack_params([p], device="cuda"),
                                             np.array([1, 0, 0]), span, rtol=1e-9, atol=1e-9, t_eval=te, return_stats=True)
    assert int(status.abs().sum()) == 0 and got.shape == (B, 7, 3)
    # Sampled envelopes make the controller reject ~40% of its steps at sample boundaries, and which side of a
    # boundary a stage lands on is decided at the 1e-16 level: CUDA and numpy follow step sequences that differ
    # in a handful of decisions (test_dopri5_smooth_case_is_step_exact shows the controller itself is exact).
    # The two runs can only agree to the controller's own global accuracy on this discontinuous RHS (~1e-6 at
    # tol 1e-9: that is also how far the ORACLE is from the tight-tolerance truth), so the check is (i) a few
    # 1e-6 between them and (ii) CUDA is as close to the truth as the oracle is.
    got = got.cpu().numpy()
    assert np.abs(got - want).max() < 5e-6
    _, truth = O.dop853_solve(om, samples, np.array([1, 0, 0]), span, te)
    err_cuda, err_oracle = np.abs(got - truth).max(), np.abs(want - truth).max()
    assert err_cuda < 2.0 * err_oracle + 1e-7, (err_cuda, err_oracle)
    assert np.abs(ns.cpu().numpy() - steps).max() <= 0.05 * steps.max()


def test_dopri5_smooth_case_is_step_exact():
    """Constant envelope over a window longer than the span: no discontinuity, so the CUDA controller must take
    exactly the oracle's accept/reject sequence and agree to rounding."""
    om, nm, ch = _models((0,), "full")
    p = dict(amp=np.array([0.6, 0.2, 0.9]), angle=np.array([0.3, -1.0, 2.0]))
    samples = O.channel_samples(ch, [O.PulseSpec("constant", "d0", 40)], [p])
    span = (0.0, 24 * DT)
    te = np.linspace(0, 24 * DT, 5)
    for tol in (1e-6, 1e-8, 1e-10):
        _, want, steps = O.dopri5_jax_solve(om, samples, np.array([1, 0, 0]), span, te, rtol=tol, atol=tol)
        got, ns, status = engine.solve_sv_dopri5(nm, [("constant", 0, 0, 40)], engine.pack_params([p], device="cuda"),
                                                 np.array([1, 0, 0]), span, rtol=tol, atol=tol, t_eval=te, return_stats=True)
        assert np.array_equal(ns.cpu().numpy(), steps) and int(status.abs().sum()) == 0
        assert np.abs(got.cpu().numpy() - want).max() < 1e-12


def test_dopri5_far_from_the_schedule_takes_the_reference_paths():
    """t / dt beyond the int range and phases beyond the branch-free sincos: the out-of-line reference routines (numpy
    floor-division index, library sincos) take over.  Nothing is played there and the frame is H0, so the state must not
    move, the error estimate is exactly zero and the controller grows the step tenfold per attempt like the oracle's."""
    om, nm, ch = _models((0,), "full")
    p = dict(amp=np.array([0.4, 0.7]), angle=0.0, sigma=8.0, beta=np.array([1.0, -2.0]))
    samples = O.channel_samples(ch, [O.PulseSpec("drag", "d0", 32)], [p])
    y0 = np.array([0.6, 0.0, 0.8j])
    span = (1.0, 1.0 + 32 * DT)  # t / dt = 4.5e9
    _, want, steps = O.dopri5_jax_solve(om, samples, y0, span, rtol=1e-8, atol=1e-8)
    got, ns, status = engine.solve_sv_dopri5(nm, [("drag", 0, 0, 32)], engine.pack_params([p], device="cuda"), y0, span,
                                             rtol=1e-8, atol=1e-8, return_stats=True)
    assert int(status.abs().sum()) == 0 and np.array_equal(ns.cpu().numpy(), steps)
    assert np.abs(got.cpu().numpy() - want).max() < 1e-14 and np.abs(got[:, -1].cpu().numpy() - y0).max() < 1e-14


def test_dopri5_mxstep_reports_failure():
    om, nm, ch = _models((0,), "full")
    p = dict(amp=0.5, angle=0.0, sigma=8.0, beta=1.0)
    out, ns, status = engine.solve_sv_dopri5(nm, [("drag", 0, 0, 32)], engine.pack_params([p], device="cuda"),
                                             np.array([1, 0, 0]), (0.0, 32 * DT), rtol=1e-10, atol=1e-10, mxstep=5,
                                             return_stats=True)
    assert int(status[0]) == 1 and bool(torch.isnan(out[0, -1].real).all())


def test_backend_solve_matches_oracle_flow():
    """The drop-in call: same circuit API as casq, results vs the oracle's restatement of casq's flow."""
    h = TransmonModel.manila(extracted_qubits=[0])
    c = ControlModel(dt=DT, channel_carrier_freqs=O.MANILA_CARRIERS)
    backend = build(BackendLibrary.B200, h, c, seed=42)
    om, _, ch = _models((0,), "full")
    gate = DragPulseGate(duration=48, amplitude=0.7)
    circuit = PulseCircuit.from_pulse_gate(gate, {"sigma": 12.0, "beta": 1.5})
    spec, par = [O.PulseSpec("drag", "d0", 48)], [dict(amp=0.7, angle=0.0, sigma=12.0, beta=1.5)]
    M = PulseBackend.ODESolverMethod
    sol = backend.solve(circuit, M.QISKIT_DYNAMICS_RK4, steps=9, run_options={"max_dt": DT / 40})
    ref = O.casq_solve(om, spec, par, method="RK4", steps=9, run_options={"max_dt": DT / 40})
    assert isinstance(sol, PulseBackend.Solution) and len(sol.states) == 9 and len(sol.counts) == 9
    np.testing.assert_allclose(sol.times, ref["times"], rtol=0, atol=0)
    got = np.array([s.data for s in sol.states])
    assert np.abs(got - ref["states"][0]).max() < TOL_STATE
    for a, b in zip(sol.populations, ref["populations"]):
        assert set(a) <= {"0", "1"} and all(abs(a.get(k, 0.0) - b.get(k, 0.0)) < 1e-9 for k in ("0", "1"))
    assert all(sum(cn.values()) == 1024 for cn in sol.counts) and sol.counts[0] == {"0": 1024}
    assert sol.shots == 1024 and len(sol.iq_data[-1]) == 1024 and len(sol.samples[-1]) == 1024
    # adaptive, default (jax) tolerances and the scipy-named method: close to the tight-tolerance truth
    truth = O.casq_solve(om, spec, par, method="DOP853", run_options={"rtol": 1e-12, "atol": 1e-12})["states"][0, -1]
    for m in (M.QISKIT_DYNAMICS_JAX_ODEINT, M.SCIPY_DOP853):
        s2 = backend.solve(circuit, m)
        assert len(s2.states) == 2 and np.abs(s2.states[-1].data - truth).max() < 1e-4  # solver accuracy at 1.4e-8
    # seeded sampling is reproducible; enum wins over run_options["method"]
    s3 = backend.solve(circuit, M.QISKIT_DYNAMICS_RK4, run_options={"max_dt": DT / 40, "method": "DOP853"})
    s4 = backend.solve(circuit, M.QISKIT_DYNAMICS_RK4, run_options={"max_dt": DT / 40})
    assert s3.counts == s4.counts and s3.samples == s4.samples
    with pytest.raises(CasqError, match="max_dt"):
        backend.solve(circuit, M.QISKIT_DYNAMICS_RK4)
    with pytest.raises(CasqError, match="LSODA"):
        backend.solve(circuit, M.SCIPY_LSODA)


def test_backend_solve_batch_rabi_golden():
    """BASELINE cfg 1 (reduced duration): Gaussian amplitude sweep vs the DOP853 truth fixture; fixed-step RK4 at
    dt/400 is within 1e-6 of it, and B = 1 reduces to solve()."""
    g = np.load(os.path.join(G, "rabi_gaussian_dop853.npz"))
    h = TransmonModel.manila(extracted_qubits=[0])
    backend = build(BackendLibrary.B200, h, ControlModel(dt=DT, channel_carrier_freqs=O.MANILA_CARRIERS))
    gate = GaussianPulseGate(duration=int(g["duration"]), amplitude=1.0)
    B = len(g["amp"])
    sol = backend.solve_batch(gate, {"sigma": np.full(B, float(g["sigma"]))}, PulseBackend.ODESolverMethod.QISKIT_DYNAMICS_JAX_ODEINT,
                              amplitude=g["amp"], run_options={"rtol": 1e-11, "atol": 1e-11, "hmax": DT / 4})
    assert np.abs(sol.states[:, -1].cpu().numpy() - g["states"][:, -1]).max() < 1e-6
    single = backend.solve(PulseCircuit.from_pulse_gate(GaussianPulseGate(int(g["duration"]), float(g["amp"][3])), {"sigma": float(g["sigma"])}),
                           PulseBackend.ODESolverMethod.QISKIT_DYNAMICS_JAX_ODEINT, run_options={"rtol": 1e-11, "atol": 1e-11, "hmax": DT / 4})
    assert np.abs(single.states[-1].data - sol.states[3, -1].cpu().numpy()).max() < 1e-13
    arr = backend.solve_batch(DragPulseGate(32, 0.3), np.array([[8.0, 1.0], [9.0, -1.0]]), PulseBackend.ODESolverMethod.QISKIT_DYNAMICS_RK4,
                              run_options={"max_dt": DT / 40})
    assert arr.batch_size == 2 and arr.states.shape == (2, 2, 3)


def test_bad_arguments_raise():
    om, nm, ch = _models((0,), "full")
    params = engine.pack_params([dict(amp=0.5, sigma=4.0)], device="cuda")
    with pytest.raises(CasqError, match="channel"):
        engine.solve_sv_rk4(nm, [("gaussian", 7, 0, 16)], params, np.array([1, 0, 0]), (0, 16 * DT), DT / 40)
    with pytest.raises(CasqError, match="max_dt"):
        engine.solve_sv_rk4(nm, [("gaussian", 0, 0, 16)], params, np.array([1, 0, 0]), (0, 16 * DT), 0.0)
    with pytest.raises(CasqError, match="increasing"):
        engine.solve_sv_rk4(nm, [("gaussian", 0, 0, 16)], params, np.array([1, 0, 0]), (0, 16 * DT), DT / 40, [2 * DT, DT])
    with pytest.raises(CasqError, match="y0"):
        engine.solve_sv_rk4(nm, [("gaussian", 0, 0, 16)], params, np.array([1, 0]), (0, 16 * DT), DT / 40)
    with pytest.raises(CasqError, match="CUDA"):
        engine.solve_sv_rk4(nm, [("gaussian", 0, 0, 16)], params.cpu(), np.array([1, 0, 0]), (0, 16 * DT), DT / 40)


def test_discretize_matches_oracle_signals():
    """casq.common.helpers.discretize: per-channel sampled signals with I/Q components (helpers.py:105-144)."""
    from casq_b200.common import discretize
    from casq_b200.gates import GaussianSquareDragPulseGate

    gate = GaussianSquareDragPulseGate(duration=40, amplitude=0.6, angle=0.4)
    sched = gate.schedule({"sigma": 5.0, "width": 12.0, "beta": 1.5}, 0)
    sigs = discretize(sched, DT, {"d0": 4.96e9, "u0": 4.84e9})
    assert [s.name for s in sigs] == ["d0", "u0"] and sigs[0].duration == 40 and sigs[0].carrier == 4.96e9
    want = O.envelope_samples("gaussian_square_drag", 40, amp=0.6, angle=0.4, sigma=5.0, width=12.0, beta=1.5)
    assert np.abs(sigs[0].signal.samples - want).max() < 1e-14
    assert np.abs(sigs[0].q_signal.samples - (want.imag - 1j * want.real)).max() < 1e-14
    assert sigs[0].i_signal.carrier_freq == pytest.approx(2 * 4.96e9) and not sigs[1].signal.samples.any()
    assert discretize(sched, DT, {"d0": 4.96e9}, carrier_frequency=1e8)[0].carrier == 1e8


def test_dopri5_late_window_and_exact_sample_boundaries():
    """(i) A pulse played 30 000 samples into the schedule: carrier phases pass 1e5 rad, where the branch-free sincos of
    the Dopri5 kernel hands over to the library one.  (ii) hmax = dt/2 with a wide tolerance makes most steps end EXACTLY
    on multiples of dt/2 accumulated in floating point: the FMA-corrected sample index must agree with numpy's
    floor-division on those boundary times (a wrong index is a 1e-2 error here, not a rounding one)."""
    om, nm, ch = _models((0,), "full")
    start, dur = 30000, 12
    p = dict(amp=np.array([0.5, 0.9]), angle=np.array([0.2, -0.7]))
    samples = O.channel_samples(ch, [O.PulseSpec("constant", "d0", dur + 4, start)], [p])
    span = (start * DT, (start + dur) * DT)
    assert 2 * np.pi * O.MANILA_CARRIERS["d0"] * span[0] > 1.0e5
    te = np.linspace(span[0], span[1], 4)
    _, want, steps = O.dopri5_jax_solve(om, samples, np.array([1, 0, 0]), span, te, rtol=1e-9, atol=1e-9)
    got, ns, status = engine.solve_sv_dopri5(nm, [("constant", 0, start, dur + 4)], engine.pack_params([p], device="cuda"),
                                             np.array([1, 0, 0]), span, rtol=1e-9, atol=1e-9, t_eval=te, return_stats=True)
    assert int(status.abs().sum()) == 0 and np.array_equal(ns.cpu().numpy(), steps)  # smooth inside the window: step-exact
    assert np.abs(got.cpu().numpy() - want).max() < 1e-10
    # (ii)
    rng = np.random.default_rng(5)
    B, dur = 4, 16
    p = _rand_params(rng, "drag", B, dur)
    samples = O.channel_samples(ch, [O.PulseSpec("drag", "d0", dur)], [p])
    span = (0.0, dur * DT)
    _, want, steps = O.dopri5_jax_solve(om, samples, np.array([1, 0, 0]), span, rtol=1e-3, atol=1e-3, hmax=DT / 2)
    got, ns, status = engine.solve_sv_dopri5(nm, [("drag", 0, 0, dur)], engine.pack_params([p], device="cuda"), np.array([1, 0, 0]),
                                             span, rtol=1e-3, atol=1e-3, hmax=DT / 2, return_stats=True)
    assert int(status.abs().sum()) == 0 and np.array_equal(ns.cpu().numpy(), steps)
    assert np.abs(got.cpu().numpy() - want).max() < 1e-11


def _small2_cases():
    """(label, h_dict, carriers, frame): models that take the two-trajectories-per-thread kernel (one driven channel,
    no RWA, nothing static left in the frame): d = 2, 3, 4 with ladder (TRI) and dense (non-TRI) couplings."""
    v = {"wq0": 2 * np.pi * 4.9e9, "om0": 6e8, "wq1": 2 * np.pi * 5.1e9, "om1": 5e8, "j": 4e7, "al": -2e9}
    car = {"d0": 4.9e9, "d1": 5.1e9}
    yield "d2", {"h_str": ["wq0/2*(I0-Z0)", "om0*X0||D0"], "qub": {"0": 2}, "vars": v}, car, "full"
    yield "d3-tri", O.MANILA, O.MANILA_CARRIERS, "full"
    yield "d3-dense", {"h_str": ["wq0*O0", "al/2*O0*O0", "-al/2*O0", "om0*X0||D0", "om0*Sp0*Sp0||D0", "om0*Sm0*Sm0||D0"],
                       "qub": {"0": 3}, "vars": v}, car, "diag"
    yield "d4-tri", {"h_str": ["wq0*O0", "al/2*O0*O0", "-al/2*O0", "om0*X0||D0"], "qub": {"0": 4}, "vars": v}, car, "full"
    yield "d4-dense", {"h_str": ["_SUM[i,0,1,wq{i}/2*(I{i}-Z{i})]", "j*Sp0*Sm1", "j*Sm0*Sp1", "om0*X0||D0"],
                       "qub": {"0": 2, "1": 2}, "vars": v}, car, "full"


@pytest.mark.parametrize("case", list(_small2_cases()), ids=lambda c: c[0])
def test_rk4_two_per_thread_kernel_is_bit_identical_to_one_per_thread(case, monkeypatch):
    """The headline kernel (sv_rk4_small2_kernel, two trajectories per thread, selected for B >= 16384) claims "same
    arithmetic, same order" as the one-trajectory kernel: assert torch.equal over a full 1e5-point batch (odd B too,
    so that the last thread owns a single trajectory), and pin both to the oracle on a handful of points."""
    _, h_dict, carriers, frame = case
    subs = [0] if h_dict is O.MANILA else None
    om, nm, ch = _models(subs, frame, None, h_dict, carriers)
    d = om.dim
    rng = np.random.default_rng(77)
    dur = 48
    span = (0.0, (dur + 1) * DT)
    slots = [("drag", ch.index("d0"), 0, dur)]
    for B in (100000, 16385):
        p = dict(amp=rng.uniform(0.02, 0.9, B), angle=rng.uniform(-np.pi, np.pi, B), sigma=rng.uniform(5, 24, B),
                 beta=rng.uniform(-4, 4, B))
        params = engine.pack_params([p], device="cuda")
        y0 = rng.normal(size=(B, d)) + 1j * rng.normal(size=(B, d))
        y0 /= np.linalg.norm(y0, axis=1, keepdims=True)
        te = np.array([0.0, 11.3, 30.0, dur + 1.0]) * DT
        out = {}
        for per in ("1", "2"):
            monkeypatch.setenv("CASQ_SMALL_TRAJ_PER_THREAD", per)
            nm.invalidate_cache()
            out[per] = engine.solve_sv_rk4(nm, slots, params, y0, span, DT / 40, te).clone()
            torch.cuda.synchronize()
        assert torch.equal(out["1"], out["2"]), case[0]
        monkeypatch.delenv("CASQ_SMALL_TRAJ_PER_THREAD")
        auto = engine.solve_sv_rk4(nm, slots, params, y0, span, DT / 40, te)  # B >= 16384: the default picks small2
        assert torch.equal(auto, out["2"])
        idx = np.array([0, 1, B // 2, B - 2, B - 1])
        samples = O.channel_samples(ch, [O.PulseSpec("drag", "d0", dur)], [{k: v[idx] for k, v in p.items()}])
        _, want = O.rk4_solve(om, samples, y0[idx], span, DT / 40, te)
        assert np.abs(auto.cpu().numpy()[idx] - want).max() < TOL_STATE
        # one piece (no t_eval): B = 1e5 takes the TIME-SLICED persistent kernel (782 blocks do not divide over 148 SMs).
        # With the stage table in shared memory (CASQ_SMALL_NO_CONST_TABLE) sliced, unsliced and one-trajectory-per-thread
        # runs must agree bit for bit, for two segment lengths.
        def run(env):
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            nm.invalidate_cache()
            r = engine.solve_sv_rk4(nm, slots, params, y0, span, DT / 40).clone()
            torch.cuda.synchronize()
            for k in env:
                monkeypatch.delenv(k)
            return r

        tab = {"CASQ_SMALL_NO_CONST_TABLE": "1"}
        runs = {"sliced4": run({**tab, "CASQ_SMALL_SLICED": "1", "CASQ_SMALL_SLICE_CHUNKS": "4"}),
                "sliced7": run({**tab, "CASQ_SMALL_SLICED": "1", "CASQ_SMALL_SLICE_CHUNKS": "7"}),
                "unsliced": run({**tab, "CASQ_SMALL_SLICED": "0"}), "one": run({"CASQ_SMALL_TRAJ_PER_THREAD": "1"}),
                "table-default": run(tab)}
        for tag in ("sliced7", "unsliced", "one", "table-default"):
            assert torch.equal(runs["sliced4"], runs[tag]), (case[0], B, tag)
        # The default is the CONSTANT-TABLE variant (offset table in the constant bank, state carried in the frame of the
        # current 32-step chunk): the same RK4 in a constant diagonal change of variables per chunk, so it agrees with the
        # table kernel to rounding (not bit for bit), with itself bit for bit whatever the slicing, and with the oracle.
        ct = {"ct-sliced4": run({"CASQ_SMALL_SLICED": "1", "CASQ_SMALL_SLICE_CHUNKS": "4"}),
              "ct-sliced7": run({"CASQ_SMALL_SLICED": "1", "CASQ_SMALL_SLICE_CHUNKS": "7"}),
              "ct-unsliced": run({"CASQ_SMALL_SLICED": "0"}), "default": run({})}
        for tag in ("ct-sliced7", "ct-unsliced", "default"):
            assert torch.equal(ct["ct-sliced4"], ct[tag]), (case[0], B, tag)
        assert float((ct["default"] - runs["sliced4"]).abs().max()) < 1e-12, (case[0], B)
        _, want1 = O.rk4_solve(om, samples, y0[idx], span, DT / 40)
        assert np.abs(ct["default"].cpu().numpy()[idx] - want1).max() < TOL_STATE


@pytest.mark.parametrize("span_samples", [(0.0, 0.6), (3.0, 20.5), (7.25, 48.0)])
def test_rk4_constant_table_variant_spans(span_samples, monkeypatch):
    """The constant-table variant of the two-trajectory kernel on spans that do not start at t = 0 (the frame of chunk 0 is
    not the rotating frame then), that end inside a chunk, and on a span of a single chunk: equal to the shared-memory-table
    kernel to rounding and to the oracle on a few points."""
    om, nm, ch = _models((0,), "full")
    rng = np.random.default_rng(5)
    B, dur = 16400, 48
    p = dict(amp=rng.uniform(0.05, 0.9, B), angle=rng.uniform(-np.pi, np.pi, B), sigma=rng.uniform(5, 20, B), beta=rng.uniform(-3, 3, B))
    params = engine.pack_params([p], device="cuda")
    y0 = rng.normal(size=(B, 3)) + 1j * rng.normal(size=(B, 3))
    y0 /= np.linalg.norm(y0, axis=1, keepdims=True)
    span = (span_samples[0] * DT, span_samples[1] * DT)
    slots = [("drag", 0, 0, dur)]
    nm.invalidate_cache()
    got = engine.solve_sv_rk4(nm, slots, params, y0, span, DT / 40).clone()
    monkeypatch.setenv("CASQ_SMALL_NO_CONST_TABLE", "1")
    nm.invalidate_cache()
    ref = engine.solve_sv_rk4(nm, slots, params, y0, span, DT / 40).clone()
    assert float((got - ref).abs().max()) < 1e-12
    assert torch.equal(got[:, 0], ref[:, 0])  # the initial state is written before any frame change
    idx = np.array([0, 1, B // 2, B - 1])
    samples = O.channel_samples(ch, [O.PulseSpec("drag", "d0", dur)], [{k: v[idx] for k, v in p.items()}])
    _, want = O.rk4_solve(om, samples, y0[idx], span, DT / 40)
    assert np.abs(got.cpu().numpy()[idx] - want).max() < TOL_STATE


def test_cfg5_dopri5_256_candidates_against_oracle_and_truth():
    """BASELINE cfg 5 settings (DRAG duration 256, amp 0.12, (sigma, beta) ~ U[16,128] x U[-4,4], hmax = dt) on 256
    candidates against the oracle's jax-odeint restatement AND a tight-tolerance DOP853 truth.

    (1) At tight tolerances (rtol = atol = 1e-10) north_star's 1e-6 bar on the state vector is ASSERTED over the first 32
        candidates (the oracle itself sits 9e-7 from a DOP853 truth there).
    (2) At cfg 5's own loose tolerances (atol 1e-6, rtol 1e-8) an accepted step may carry a local error of ~1e-6 and
        the controller rejects ~3 % of its steps at sample boundaries; which side of a boundary a stage time lands on is
        decided in the last bits of h, so CUDA and numpy follow different step sequences for some candidates and then
        differ by the METHOD's own global error at this tolerance (the oracle is 4e-5 .. 2e-4 from the truth here; CUDA vs
        oracle measured on a B200: 5.5e-5 worst of 256, median 1e-11).  No implementation reproduces another to 1e-6 there -- the reference's XLA build would not either -- so
        the asserted bound is the stated looser one, 2e-4, together with: CUDA is as close to the truth as the oracle is,
        and most candidates (the median) agree with the oracle to 1e-9.  (Equal step COUNTS do not imply equal step sequences:
        an accept/reject flip at one boundary is usually compensated at the next.)"""
    from _oracle_pool import cfg5_oracle

    om, nm, ch = _models((0,), "full")
    rng = np.random.default_rng(20260101)
    B = 256
    sigma, beta = rng.uniform(16, 128, 10000)[:B], rng.uniform(-4, 4, 10000)[:B]
    p = dict(amp=0.12, angle=0.0, sigma=sigma, beta=beta)
    params = engine.pack_params([p], device="cuda")
    span = (0.0, 256 * DT)
    y0 = np.array([1, 0, 0])
    # (1) tight tolerances
    want_t, _ = cfg5_oracle(sigma[:32], beta[:32], tol=(1e-10, 1e-10))
    got_t, _, status = engine.solve_sv_dopri5(nm, [("drag", 0, 0, 256)], params[:, :, :32].contiguous(), y0, span, rtol=1e-10,
                                              atol=1e-10, hmax=DT, return_stats=True)
    assert int(status.abs().sum()) == 0
    assert np.abs(got_t[:, -1].cpu().numpy() - want_t).max() < 1e-6
    # (2) cfg 5's own tolerances
    want, steps = cfg5_oracle(sigma, beta)
    truth, _ = cfg5_oracle(sigma, beta, truth=True)
    got, ns, status = engine.solve_sv_dopri5(nm, [("drag", 0, 0, 256)], params, y0, span, rtol=1e-8, atol=1e-6, hmax=DT,
                                             return_stats=True)
    assert int(status.abs().sum()) == 0
    got, ns = got[:, -1].cpu().numpy(), ns.cpu().numpy()
    err = np.abs(got - want).max(axis=1)
    assert err.max() < 2e-4, (err.max(), int(err.argmax()))
    assert np.median(err) < 1e-9
    assert (err < 1e-9).sum() >= B // 2  # most candidates follow the oracle's step sequence exactly
    e_cuda, e_oracle = np.abs(got - truth).max(axis=1), np.abs(want - truth).max(axis=1)
    assert e_cuda.max() < 2.0 * e_oracle.max() + 1e-7 and np.median(e_cuda) < 2.0 * np.median(e_oracle) + 1e-9
    assert np.abs(ns - steps).max() <= 0.05 * steps.max()
