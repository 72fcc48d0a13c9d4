"""Oracle: integrators of y' = G_F(t) y  (and the Lindblad equation) for batches of pulse points.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED: restates qiskit-dynamics 0.4.1
``solve_lmde``/``solve_ode`` back ends selected by casq's ``ODESolverMethod``
(src/casq/backends/pulse_backend.py:60-71; call site src/casq/backends/qiskit/
qiskit_pulse_backend.py:158-162) -- SURVEY Appendix A.5:
  * "RK4"/"jax_RK4": fixed_step_solver_template + get_fixed_step_sizes + RK4 take_step
  * scipy names: scipy.integrate.solve_ivp on the flattened complex state
  * "jax_odeint": jax 0.4.6 jax.experimental.ode.odeint (Dopri5, FSAL, 4th-order dense output)
  * exponential-midpoint propagators (upstream "scipy_expm"; not in casq's enum, spec for cfg 3)
  * LindbladModel.evaluate_rhs (Appendix A.4 last bullet) for cfg 4
All solvers are vectorised over a leading batch axis so tests can check many points at once;
one point is B=1.  States are integrated in the frame basis of the rotating frame and returned
"in the frame, out of the frame basis", exactly what Solver.solve hands to casq's
get_experiment_result (src/casq/backends/qiskit/helpers.py:98-109).
"""
from __future__ import annotations

import numpy as np
from scipy.integrate import solve_ivp
from scipy.linalg import expm

from .model import OracleModel
from .pulses import floor_divide_index


def merge_t_args(t_span, t_eval):
    t_span = np.asarray(t_span, dtype=float)
    if t_eval is None:
        return t_span.copy()
    t_eval = np.asarray(t_eval, dtype=float)
    out = t_eval
    if t_eval[0] != t_span[0]:
        out = np.append(t_span[0], out)
    if t_eval[-1] != t_span[1]:
        out = np.append(out, t_span[1])
    return out


def fixed_step_sizes(t_span, t_eval, max_dt):
    """get_fixed_step_sizes: cut the span at t_eval; per piece n = int(|D/max_dt|) corrected upward."""
    pts = merge_t_args(t_span, t_eval)
    deltas = np.diff(pts)
    t_list = pts[:-1]
    n_steps = np.abs(deltas / max_dt).astype(int)
    for i, (dlt, n) in enumerate(zip(deltas, n_steps)):
        if n == 0:
            n_steps[i] = 1
        elif np.abs(dlt / n) / max_dt > 1 + 1e-15:
            n_steps[i] = n + 1
    h_list = deltas / n_steps
    return t_list, h_list, n_steps


def _z_of_t(model: OracleModel, samples, t):
    """z[B,K] at time t: samples[b,k,idx(t)] * exp(i 2 pi nu_k t), zero outside the window."""
    N = samples.shape[-1]
    idx = int(floor_divide_index(t, model.dt, N))
    if idx < 0 or idx >= N:
        return np.zeros(samples.shape[:2], dtype=complex)
    return samples[:, :, idx] * np.exp(1j * (2 * np.pi * model.nu) * t)


def _rhs_sv(model, samples, t, y):
    G = model.generator_fb(t, _z_of_t(model, samples, t))  # [B,d,d]
    return np.einsum("bij,bj->bi", G, y)


def _into_fb(model, y0, B):
    y0 = np.asarray(y0, dtype=complex)
    if y0.ndim == 1:
        y0 = np.broadcast_to(y0, (B, len(y0)))
    return y0 @ model.frame_basis.conj()  # U^+ y  (row-vector form)


def _out_of_fb(model, y):
    return y @ model.frame_basis.T  # U y


def rk4_solve(model: OracleModel, samples, y0, t_span, max_dt, t_eval=None):
    """Fixed-step RK4 as qiskit-dynamics RK4_solver.  samples [B,K,N]; returns (t [T], y [B,T,d])."""
    B = samples.shape[0]
    t_list, h_list, n_list = fixed_step_sizes(t_span, t_eval, max_dt)
    y = _into_fb(model, y0, B).copy()
    ys = [y]
    for t0, h, n in zip(t_list, h_list, n_list):
        t = t0
        h2 = 0.5 * h
        for _ in range(int(n)):
            tp = t + h2
            k1 = _rhs_sv(model, samples, t, y)
            k2 = _rhs_sv(model, samples, tp, y + h2 * k1)
            k3 = _rhs_sv(model, samples, tp, y + h2 * k2)
            k4 = _rhs_sv(model, samples, t + h, y + h * k3)
            y = y + (1.0 / 6) * h * (k1 + 2 * k2 + 2 * k3 + k4)
            t = t + h
        ys.append(y)
    ts = merge_t_args(t_span, t_eval)
    out = _out_of_fb(model, np.stack(ys, axis=1))
    if t_eval is not None:  # trim to t_eval (upstream trim_t_results)
        keep = np.isin(ts, np.asarray(t_eval, dtype=float))
        ts, out = ts[keep], out[:, keep]
    return ts, out


def dop853_solve(model: OracleModel, samples, y0, t_span, t_eval=None, rtol=1e-12, atol=1e-12, method="DOP853",
                 max_step=np.inf):
    """scipy solve_ivp per point ("truth" integrator).  Returns (t, y [B,T,d]); t_eval None -> endpoints."""
    B = samples.shape[0]
    y0fb = _into_fb(model, y0, B)
    outs = []
    te = np.asarray(t_span, dtype=float) if t_eval is None else np.asarray(t_eval, dtype=float)
    for b in range(B):
        sb = samples[b : b + 1]
        # the envelope is piecewise constant: integrate sample by sample so the adaptive
        # controller never straddles a discontinuity (tight-tolerance truth, not a casq option)
        edges = np.unique(np.concatenate([te, np.arange(0, samples.shape[-1] + 1) * model.dt]))
        edges = edges[(edges >= t_span[0]) & (edges <= t_span[1])]
        y = y0fb[b].astype(complex)
        res = {edges[0]: y}
        for a, c in zip(edges[:-1], edges[1:]):
            mid = 0.5 * (a + c)
            z = _z_of_t(model, sb, mid) * np.exp(-1j * (2 * np.pi * model.nu) * mid)  # bare env on this piece

            def rhs_piece(t, yy, z=z):
                zz = z * np.exp(1j * (2 * np.pi * model.nu) * t)
                return model.generator_fb(t, zz)[0] @ yy

            sol = solve_ivp(rhs_piece, (a, c), y, method=method, rtol=rtol, atol=atol, max_step=max_step)
            y = sol.y[:, -1]
            res[c] = y
        outs.append(np.stack([res[t] for t in te]))
    return te, _out_of_fb(model, np.stack(outs))


# ---------------------------------------------------------------- jax.experimental.ode.odeint (Dopri5)
_DP_ALPHA = np.array([1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0, 0.0])
_DP_BETA = np.array([
    [1 / 5, 0, 0, 0, 0, 0, 0],
    [3 / 40, 9 / 40, 0, 0, 0, 0, 0],
    [44 / 45, -56 / 15, 32 / 9, 0, 0, 0, 0],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729, 0, 0, 0],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656, 0, 0],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0],
])
_DP_CSOL = np.array([35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0])
_DP_CERR = np.array([35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
                     -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1.0 / 60.0])
_DP_CMID = np.array([6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
                     187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2])


def _dopri5_one(func, y0, ts, rtol, atol, mxstep, hmax):
    """One trajectory, jax odeint control flow.  func(t, y).  Returns (ys [T,d], n_attempted, n_accepted)."""

    def rk_step(y, f, t, dt):
        k = np.zeros((7, len(y)), dtype=complex)
        k[0] = f
        for i in range(1, 7):
            ti = t + dt * _DP_ALPHA[i - 1]
            yi = y + dt * (_DP_BETA[i - 1, :] @ k)
            k[i] = func(ti, yi)
        y1 = dt * (_DP_CSOL @ k) + y
        err = dt * (_DP_CERR @ k)
        return y1, k[6], err, k

    def err_ratio(err, ya, yb):
        tol = atol + rtol * np.maximum(np.abs(ya), np.abs(yb))
        r = err / tol
        return np.sqrt(np.mean(r.real**2 + r.imag**2))

    def opt_step(last, ratio):
        if ratio == 0:
            return last * 10.0
        dfac = 1.0 if ratio < 1 else 0.2
        return last * min(10.0, max(ratio ** (-1.0 / 5.0) * 0.9, dfac))

    def fit(ya, yb, k, dt):
        ymid = ya + dt * (_DP_CMID @ k)
        dy0, dy1 = k[0], k[6]
        a = -2.0 * dt * dy0 + 2.0 * dt * dy1 - 8.0 * ya - 8.0 * yb + 16.0 * ymid
        b = 5.0 * dt * dy0 - 3.0 * dt * dy1 + 18.0 * ya + 14.0 * yb - 32.0 * ymid
        c = -4.0 * dt * dy0 + dt * dy1 - 11.0 * ya - 5.0 * yb + 16.0 * ymid
        return [a, b, c, dt * dy0, ya]

    # initial step (Hairer II.4), order 4
    f0 = func(ts[0], y0)
    scale = atol + np.abs(y0) * rtol
    d0 = np.linalg.norm(y0 / scale)
    d1 = np.linalg.norm(f0 / scale)
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    f1 = func(ts[0] + h0, y0 + h0 * f0)
    d2 = np.linalg.norm((f1 - f0) / scale) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = max(1e-6, h0 * 1e-3)
    else:
        h1 = (0.01 / (d1 + d2)) ** (1.0 / 5.0)
    dt = float(np.clip(min(100.0 * h0, h1), 0.0, hmax))

    y, f, t, last_t = y0.copy(), f0, ts[0], ts[0]
    coeff = [y0.copy() for _ in range(5)]
    out = [y0.copy()]
    n_att = n_acc = 0
    for target in ts[1:]:
        i = 0
        while t < target and i < mxstep and dt > 0:
            y1, f1n, err, k = rk_step(y, f, t, dt)
            ratio = err_ratio(err, y, y1)
            new_coeff = fit(y, y1, k, dt)
            new_dt = float(np.clip(opt_step(dt, ratio), 0.0, hmax))
            n_att += 1
            i += 1
            if ratio <= 1.0:
                last_t, t, y, f, coeff = t, t + dt, y1, f1n, new_coeff
                n_acc += 1
            dt = new_dt
        rel = (target - last_t) / (t - last_t)
        a, b, c, d, e = coeff
        out.append((((a * rel + b) * rel + c) * rel + d) * rel + e)
    return np.stack(out), n_att, n_acc


def dopri5_jax_solve(model: OracleModel, samples, y0, t_span, t_eval=None, rtol=1.4e-8, atol=1.4e-8,
                     mxstep=np.inf, hmax=np.inf):
    """jax_odeint path.  Note jnp.max(d1+d2) on a scalar is the scalar: h1=(0.01/(d1+d2))^(1/5).

    Returns (t [T], y [B,T,d], steps [B,2] = attempted, accepted).
    """
    B = samples.shape[0]
    ts = merge_t_args(t_span, t_eval)
    y0fb = _into_fb(model, y0, B)
    ys, steps = [], []
    for b in range(B):
        sb = samples[b : b + 1]
        yb, na, nc = _dopri5_one(lambda t, y: _rhs_sv(model, sb, t, y[None])[0], y0fb[b].astype(complex), ts,
                                 rtol, atol, mxstep, hmax)
        ys.append(yb)
        steps.append((na, nc))
    out = _out_of_fb(model, np.stack(ys))
    if t_eval is not None:
        keep = np.isin(ts, np.asarray(t_eval, dtype=float))
        ts, out = ts[keep], out[:, keep]
    return ts, out, np.array(steps)


# ---------------------------------------------------------------- exponential midpoint propagators
def expm_midpoint_propagators(model: OracleModel, samples, t_span, max_dt):
    """U = prod_n expm(h G_F(t_n + h/2)) (upstream scipy_expm_solver on y0 = identity).

    Returned in the frame, out of the frame basis: U_rf = V U_fb V^+.  [B,d,d]
    """
    B = samples.shape[0]
    d = model.dim
    t_list, h_list, n_list = fixed_step_sizes(t_span, None, max_dt)
    U = np.broadcast_to(np.eye(d, dtype=complex), (B, d, d)).copy()
    for t0, h, n in zip(t_list, h_list, n_list):
        t = t0
        for _ in range(int(n)):
            G = model.generator_fb(t + 0.5 * h, _z_of_t(model, samples, t + 0.5 * h))
            E = np.stack([expm(h * G[b]) for b in range(B)])
            U = E @ U
            t = t + h
    V = model.frame_basis
    return V @ U @ V.conj().T


# ---------------------------------------------------------------- Lindblad
def lindblad_rhs_dense(model: OracleModel, t, z, rho_fb):
    """LindbladModel.evaluate_rhs restated in closed form (frame F=-iH_f, frame basis):

    rho_F' = -i[H_F(t), rho_F] + sum_j (L_j(t) rho_F L_j(t)^+ - 1/2 {L_j(t)^+ L_j(t), rho_F}),
    with every operator element carrying its Bohr phase, L_j(t)_ab = e^{i(e_a-e_b)t} L_ab.
    (RWA, if any, touches Hamiltonian operators only.)  rho_fb [B,d,d].
    """
    H = model.hamiltonian_fb(t, z)  # [B,d,d]
    ph = np.exp(1j * model.bohr * t)
    # (L(t)^+ L(t))_ab = e^{i(e_a-e_b)t} (L^+ L)_ab, so the anticommutator part needs one static matrix D = sum_j L_j^+ L_j:
    #   rho' = A rho + rho A^+ + sum_j L_j(t) rho L_j(t)^+,   A = -i H_F(t) - 1/2 e^{i bohr t} o D
    cache = getattr(model, "_lindblad_cache", None)
    if cache is None:
        D = sum(L.conj().T @ L for L in model.diss_fb)
        diag_w = np.zeros_like(D)  # diagonal L_j: L rho L^+ is an elementwise product with l l^+
        dense = []
        for L in model.diss_fb:
            if np.count_nonzero(L - np.diag(np.diag(L))) == 0:
                l = np.diag(L)
                diag_w = diag_w + np.outer(l, l.conj())
            else:
                dense.append(L)
        cache = model._lindblad_cache = (D, diag_w, dense)
    D, diag_w, dense = cache
    A = -1j * H - 0.5 * (ph * D)
    out = A @ rho_fb + rho_fb @ np.conj(np.swapaxes(A, -1, -2)) + diag_w * rho_fb
    for L in dense:
        Lt = ph * L
        out = out + Lt @ rho_fb @ Lt.conj().T
    return out


def lindblad_rk4_solve(model: OracleModel, samples, rho0, t_span, max_dt, t_eval=None):
    """Fixed-step RK4 on the density matrix (same stepping rule as rk4_solve).  Returns (t, rho [B,T,d,d])."""
    B = samples.shape[0]
    V = model.frame_basis
    rho0 = np.asarray(rho0, dtype=complex)
    if rho0.ndim == 1:
        rho0 = np.outer(rho0, rho0.conj())
    rho = np.broadcast_to(V.conj().T @ rho0 @ V, (B,) + rho0.shape).copy()
    t_list, h_list, n_list = fixed_step_sizes(t_span, t_eval, max_dt)
    rhos = [rho]

    def f(t, r):
        return lindblad_rhs_dense(model, t, _z_of_t(model, samples, t), r)

    for t0, h, n in zip(t_list, h_list, n_list):
        t = t0
        h2 = 0.5 * h
        for _ in range(int(n)):
            k1 = f(t, rho)
            k2 = f(t + h2, rho + h2 * k1)
            k3 = f(t + h2, rho + h2 * k2)
            k4 = f(t + h, rho + h * k3)
            rho = rho + (1.0 / 6) * h * (k1 + 2 * k2 + 2 * k3 + k4)
            t = t + h
        rhos.append(rho)
    ts = merge_t_args(t_span, t_eval)
    out = np.stack(rhos, axis=1)
    out = V @ out @ V.conj().T
    if t_eval is not None:
        keep = np.isin(ts, np.asarray(t_eval, dtype=float))
        ts, out = ts[keep], out[:, keep]
    return ts, out
