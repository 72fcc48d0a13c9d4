"""Oracle: casq's result post-processing and the end-to-end single-point ``solve`` flow.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED (upstream helpers restated):
  * src/casq/backends/qiskit/helpers.py:98-121  frame removal, dressed basis, normalisation
  * src/casq/backends/qiskit/helpers.py:129-141 populations with levels>1 folded to "1", sampling, counts
  * src/casq/backends/qiskit/dynamics_backend_patch.py:180-184  steps -> t_eval = linspace(t0, t1, steps)
  * src/casq/backends/qiskit/qiskit_pulse_backend.py:147-162    option packing of ``solve``
"""
from __future__ import annotations

import numpy as np

from .model import OracleModel
from .pulses import PulseSpec, channel_samples
from . import solvers


def postprocess_statevectors(model: OracleModel, times, y, normalize=True):
    """y [B,T,d] (in frame, lab basis) -> lab-frame, dressed-basis, normalised states [B,T,d]."""
    U = model.frame_basis
    _, V = model.dressed()
    times = np.asarray(times, dtype=float)
    y_fb = y @ U.conj()  # U^+ y
    y_fb = y_fb * np.exp(-1j * model.frame_evals[None, None, :] * times[None, :, None])
    y_lab = y_fb @ U.T
    out = y_lab @ V.conj()  # V^+ y
    if normalize:
        out = out / np.linalg.norm(out, axis=-1, keepdims=True)
    return out


def postprocess_density_matrices(model: OracleModel, times, rho, normalize=True):
    """rho [B,T,d,d] (in frame, lab basis) -> lab-frame dressed-basis trace-normalised."""
    U = model.frame_basis
    _, V = model.dressed()
    times = np.asarray(times, dtype=float)
    r = U.conj().T @ rho @ U
    ph = np.exp(-1j * model.frame_evals[None, :] * times[:, None])  # [T,d]
    r = ph[None, :, :, None] * r * ph.conj()[None, :, None, :]
    r = U @ r @ U.conj().T
    r = V.conj().T @ r @ V
    if normalize:
        r = r / np.trace(r, axis1=-2, axis2=-1)[..., None, None]
    return r


def level_probabilities(states):
    """|amplitude|^2 for state vectors [..., d] or the diagonal for density matrices [..., d, d]."""
    states = np.asarray(states)
    return np.abs(states) ** 2


def populations_dict(probs, subsystem_dims=None, measured=(0,)):
    """_get_memory_slot_probabilities with max_outcome_value=1 for ONE probability vector.

    casq never forwards subsystem_dims (qiskit_pulse_backend.py:184-188) so the whole system is one
    subsystem of dim d: outcome "0" = level 0, every other level folds into "1" (helpers.py:129-134).
    With subsystem_dims given, qubit q's level is digit q (little-endian) and each is clipped to 1.
    """
    probs = np.asarray(probs, dtype=float)
    d = len(probs)
    dims = [d] if subsystem_dims is None else list(subsystem_dims)
    out: dict = {}
    for idx, p in enumerate(probs):
        levels, rem = [], idx
        for dim in dims:
            levels.append(rem % dim)
            rem //= dim
        # memory slot i <- measured subsystem i; bit-string is big-endian (slot 0 right-most)
        key = "".join(str(min(levels[q], 1)) for q in reversed(list(measured)))
        out[key] = out.get(key, 0.0) + float(p)
    return out


def sample_counts(pop: dict, shots: int, seed=None):
    """_sample_probability_dict + _get_counts_from_samples."""
    keys = list(pop.keys())
    p = np.array([pop[k] for k in keys], dtype=float)
    rng = np.random.default_rng(seed)
    samples = rng.choice(keys, size=shots, p=p / p.sum())
    vals, cnts = np.unique(samples, return_counts=True)
    return samples, {str(v): int(c) for v, c in zip(vals, cnts)}


def casq_solve(model: OracleModel, specs, params, method="RK4", steps=None, run_options=None, initial_state=None,
               t_final_samples=None):
    """QiskitPulseBackend.solve restated for one circuit (one or more Play instructions), or a batch.

    Returns dict(times, states [B,T,d], populations list-of-dicts for b=0, raw [B,T,d]).
    t_span = [0, t_acquire*dt]; the default `measure` acquire starts when the gate ends, so
    t_acquire = schedule duration (Appendix A.6, flagged "(verify)" in SURVEY).
    """
    run_options = dict(run_options or {})
    run_options.pop("method", None)
    samples = channel_samples(model.channels, specs, params)
    n_total = samples.shape[-1] if t_final_samples is None else t_final_samples
    t_span = (0.0, n_total * model.dt)
    t_eval = None if steps is None else np.linspace(t_span[0], t_span[1], steps)
    _, V = model.dressed()
    y0 = V[:, 0] if initial_state is None else np.asarray(initial_state, dtype=complex)
    if method in ("RK4", "jax_RK4"):
        ts, y = solvers.rk4_solve(model, samples, y0, t_span, run_options["max_dt"], t_eval)
    elif method == "jax_odeint":
        ts, y, _ = solvers.dopri5_jax_solve(model, samples, y0, t_span, t_eval, **run_options)
    else:
        ts, y = solvers.dop853_solve(model, samples, y0, t_span, t_eval, method=method,
                                     rtol=run_options.get("rtol", 1e-3), atol=run_options.get("atol", 1e-6))
    states = postprocess_statevectors(model, ts, y)
    pops = [populations_dict(np.abs(states[0, i]) ** 2) for i in range(states.shape[1])]
    return {"times": ts, "states": states, "populations": pops, "raw": y}
