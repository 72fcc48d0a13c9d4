"""Edge cases of the C-ABI solves on the GPU: empty batches, single output point, one-step pieces, chunk
boundaries of the staged stage table, free evolution, batched initial states for the adaptive solver."""
import numpy as np
import pytest
import torch

import oracle as O
from casq_b200 import engine
from casq_b200._native import NativeModel

pytestmark = pytest.mark.gpu
DT = O.MANILA_DT


@pytest.fixture(scope="module")
def q0():
    st, ops, ch, _ = O.parse_hamiltonian_dict(O.MANILA, [0])
    om = O.OracleModel(st, ops, ch, O.MANILA_CARRIERS, DT, rotating_frame=st)
    nm = NativeModel(st, ops, [O.MANILA_CARRIERS[c] for c in ch], DT, rotating_frame=st)
    return om, nm, ch


def test_empty_batches(q0):
    om, nm, ch = q0
    empty = torch.zeros((1, 6, 0), dtype=torch.float64, device="cuda")
    slots = [("drag", 0, 0, 16)]
    assert engine.solve_sv_rk4(nm, slots, empty, np.array([1, 0, 0]), (0, 16 * DT), DT / 10).shape == (0, 2, 3)
    assert engine.solve_sv_dopri5(nm, slots, empty, np.array([1, 0, 0]), (0, 16 * DT)).shape == (0, 2, 3)
    assert engine.propagators_expm(nm, slots, empty, (0, 16 * DT), DT).shape == (0, 3, 3)


@pytest.mark.parametrize("n_steps_per_piece", [1, 31, 32, 33, 65])
def test_rk4_chunk_boundaries_and_single_output(q0, n_steps_per_piece):
    """Pieces of 1, chunk-1, chunk, chunk+1 and 2*chunk+1 steps (the table is staged in 32-step chunks)."""
    om, nm, ch = q0
    rng = np.random.default_rng(n_steps_per_piece)
    B, dur = 70, 12
    p = dict(amp=rng.uniform(0.1, 0.9, B), angle=rng.uniform(-1, 1, B), sigma=rng.uniform(2, 5, B), beta=rng.uniform(-2, 2, B))
    samples = O.channel_samples(ch, [O.PulseSpec("drag", "d0", dur)], [p])
    te = np.array([0.0, 2.5, 5.0, 7.5]) * DT
    max_dt = 2.5 * DT / n_steps_per_piece * 1.0000001
    span = (0.0, dur * DT)
    _, want = O.rk4_solve(om, samples, np.array([1, 0, 0]), span, max_dt, te)
    got = engine.solve_sv_rk4(nm, [("drag", 0, 0, dur)], engine.pack_params([p], device="cuda"), np.array([1, 0, 0]), span, max_dt, te)
    assert got.shape == (B, 4, 3) and np.abs(got.cpu().numpy() - want).max() < 1e-9
    one = engine.solve_sv_rk4(nm, [("drag", 0, 0, dur)], engine.pack_params([p], device="cuda"), np.array([1, 0, 0]), span, max_dt,
                              np.array([7.5 * DT]))
    # a different t_eval is a different time grid (t restarts at every output point), and with a sampled envelope
    # stages that land on a sample boundary can pick another sample: compare against the oracle on the SAME grid
    _, want1 = O.rk4_solve(om, samples, np.array([1, 0, 0]), span, max_dt, np.array([7.5 * DT]))
    assert one.shape == (B, 1, 3) and np.abs(one.cpu().numpy() - want1).max() < 1e-9


def test_free_evolution_is_identity_in_the_h0_frame(q0):
    om, nm, ch = q0
    params = torch.zeros((1, 6, 3), dtype=torch.float64, device="cuda")
    y0 = np.array([0.6, 0.0, 0.8j])
    out = engine.solve_sv_rk4(nm, [], params, y0, (0.0, 40 * DT), DT / 10)
    assert np.abs(out.cpu().numpy()[:, -1] - y0).max() < 1e-15
    out = engine.solve_sv_dopri5(nm, [], params, y0, (0.0, 40 * DT))
    assert np.abs(out.cpu().numpy()[:, -1] - y0).max() < 1e-15


def test_dopri5_batched_initial_states_and_window_end(q0):
    """t_span longer than the pulse window (signal is 0 after the last sample), y0 per point."""
    om, nm, ch = q0
    rng = np.random.default_rng(2)
    B, dur = 9, 10
    p = dict(amp=rng.uniform(0.2, 0.8, B), angle=0.0, sigma=3.0)
    samples = O.channel_samples(ch, [O.PulseSpec("gaussian", "d0", dur)], [p])
    y0 = rng.normal(size=(B, 3)) + 1j * rng.normal(size=(B, 3))
    y0 /= np.linalg.norm(y0, axis=1, keepdims=True)
    span = (0.0, 14 * DT)
    _, want, _ = O.dopri5_jax_solve(om, samples, y0, span, rtol=1e-8, atol=1e-8)
    got = engine.solve_sv_dopri5(nm, [("gaussian", 0, 0, dur)], engine.pack_params([p], device="cuda"), y0, span, rtol=1e-8, atol=1e-8)
    assert np.abs(got.cpu().numpy() - want).max() < 2e-6
    _, want4 = O.rk4_solve(om, samples, y0, span, DT / 16)
    got4 = engine.solve_sv_rk4(nm, [("gaussian", 0, 0, dur)], engine.pack_params([p], device="cuda"), y0, span, DT / 16)
    assert np.abs(got4.cpu().numpy() - want4).max() < 1e-9


def test_max_slots_and_two_channels_d3(q0):
    """Four Play instructions on two channels (d0 and the cross-drive u0 of the same transmon)."""
    om, nm, ch = q0
    rng = np.random.default_rng(6)
    B = 11
    specs = [O.PulseSpec("gaussian", "d0", 8, 0), O.PulseSpec("constant", "u0", 6, 2), O.PulseSpec("drag", "d0", 8, 10),
             O.PulseSpec("gaussian_square", "u0", 10, 9)]
    ps = [dict(amp=rng.uniform(0.1, 0.5, B), angle=0.1, sigma=2.0), dict(amp=rng.uniform(0.1, 0.5, B), angle=-0.4),
          dict(amp=rng.uniform(0.1, 0.5, B), angle=0.0, sigma=2.5, beta=0.7), dict(amp=rng.uniform(0.1, 0.5, B), angle=0.9, sigma=1.5, width=4.0)]
    samples = O.channel_samples(ch, specs, ps)
    slots = [(s.kind, ch.index(s.channel), s.start, s.duration) for s in specs]
    span = (0.0, 19 * DT)
    _, want = O.rk4_solve(om, samples, np.array([1, 0, 0]), span, DT / 40)
    got = engine.solve_sv_rk4(nm, slots, engine.pack_params(ps, device="cuda"), np.array([1, 0, 0]), span, DT / 40)
    assert np.abs(got.cpu().numpy() - want).max() < 1e-9
    _, wantd, _ = O.dopri5_jax_solve(om, samples[:3], np.array([1, 0, 0]), span, rtol=1e-9, atol=1e-9)
    gotd = engine.solve_sv_dopri5(nm, slots, engine.pack_params(ps, device="cuda"), np.array([1, 0, 0]), span, rtol=1e-9, atol=1e-9)
    assert np.abs(gotd.cpu().numpy()[:3] - wantd).max() < 5e-6


def test_envelope_table_windows_give_identical_results(monkeypatch):
    """Dopri5 (both kernels) and the propagator path look the envelope up in a per-solve table that is built for windows of
    the batch (bounded memory for huge sweeps).  Forcing 3-point windows must not change a single bit."""
    st, ops, ch, _ = O.parse_hamiltonian_dict(O.MANILA, [0])
    nm = NativeModel(st, ops, [O.MANILA_CARRIERS[c] for c in ch], DT, rotating_frame=st)
    rng = np.random.default_rng(12)
    B, dur = 11, 16
    p = dict(amp=rng.uniform(0.1, 0.9, B), angle=rng.uniform(-1, 1, B), sigma=4.0, beta=rng.uniform(-2, 2, B))
    pk = engine.pack_params([p], device="cuda")
    y0 = np.array([1, 0, 0])

    v = O.MANILA["vars"]
    hd = O.transmon_hamiltonian_dict({0: (v["wq0"], v["delta0"], v["omegad0"]), 1: (v["wq1"], v["delta1"], v["omegad1"])},
                                     {(0, 1): v["jq0q1"]})
    st2, ops2, ch2, _ = O.parse_hamiltonian_dict(hd, None)
    nm2 = NativeModel(st2, ops2, [O.MANILA_CARRIERS[c] for c in ch2], DT, rotating_frame=st2)  # d = 9: warp-per-trajectory Dopri5
    y02 = np.zeros(9, dtype=complex)
    y02[0] = 1

    def run():
        y, ns, stt = engine.solve_sv_dopri5(nm, [("drag", 0, 0, dur)], pk, y0, (0.0, dur * DT), rtol=1e-8, atol=1e-8, return_stats=True)
        U = engine.propagators_expm(nm, [("drag", 0, 0, dur)], pk, (0.0, dur * DT), DT / 2)
        y2, ns2, _ = engine.solve_sv_dopri5(nm2, [("drag", ch2.index("d0"), 0, dur)], pk, y02, (0.0, dur * DT), rtol=1e-8, atol=1e-8,
                                            return_stats=True)
        return y.clone(), ns.clone(), stt.clone(), U.clone(), y2.clone(), ns2.clone()

    monkeypatch.delenv("CASQ_ENV_TABLE_MAX_KB", raising=False)
    ref = run()
    monkeypatch.setenv("CASQ_ENV_TABLE_MAX_KB", "1")  # 1 kB / (16 samples x 16 B) = 4 points per window
    got = run()
    for a, b in zip(ref, got):
        assert torch.equal(a, b)
