"""B200 pulse backend: casq's PulseBackend API on top of the sm_100a engine.

Plays the role QiskitPulseBackend plays in casq (src/casq/backends/qiskit/qiskit_pulse_backend.py:45-189):
``_get_native_backend`` builds the engine model where casq builds ``Solver`` + ``DynamicsBackendPatch``
(:166-189); ``solve`` packs options, runs ONE circuit and converts the result (:119-164) with the
post-processing semantics of ``get_experiment_result`` (src/casq/backends/qiskit/helpers.py:98-162).
``solve_batch`` is the additive API: the same path for B pulse-parameter points in one launch (casq has no
sweep API; a sweep there is a Python loop of ``solve`` calls).
"""
from __future__ import annotations

import logging
from dataclasses import dataclass
from datetime import datetime
from typing import Any, Optional, Sequence, Union

import numpy as np
import torch

from ... import engine
from ..._native import NativeModel
from ...common.exceptions import CasqError
from ...gates.pulse_circuit import PulseCircuit
from ...gates.pulse_gate import PulseGate
from ...models.control_model import ControlModel
from ...models.hamiltonian_model import HamiltonianModel
from ...pulse import Play, Schedule, SetFrequency, SetPhase, ShiftFrequency, ShiftPhase
from ...quantum_info import DensityMatrix, Statevector, populations_by_qubit
from ..pulse_backend import PulseBackend

_log = logging.getLogger("casq_b200")
_FIXED = ("RK4", "jax_RK4")
_SCIPY = ("BDF", "DOP853", "Radau", "RK23", "RK45")


@dataclass
class BatchSolution:
    """Result of ``solve_batch``: device tensors for B pulse points at T time points.

    states are lab-frame, dressed-basis, normalised (helpers.py:100-109); populations[..., 0/1] are casq's
    {"0","1"} with every level >= 1 folded into "1" (helpers.py:129-134); probabilities are per level.
    """

    times: np.ndarray
    states: torch.Tensor
    probabilities: torch.Tensor
    populations: torch.Tensor
    raw: Optional[torch.Tensor] = None
    nsteps: Optional[torch.Tensor] = None
    status: Optional[torch.Tensor] = None

    @property
    def batch_size(self) -> int:
        return int(self.states.shape[0])


@dataclass
class DurationSweepSolution:
    """Result of ``solve_batch(duration=[...])``: points of different gate duration have different time grids, so ``times``
    is per point: times [B, T], states [B, T, d], probabilities [B, T, d], populations [B, T, 2]."""

    times: np.ndarray
    states: torch.Tensor
    probabilities: torch.Tensor
    populations: torch.Tensor
    durations: np.ndarray

    @property
    def batch_size(self) -> int:
        return int(self.states.shape[0])

    @staticmethod
    def assemble(B: int, parts: dict) -> "DurationSweepSolution":
        first = next(iter(parts.values()))[1]
        T, d, dev = first.states.shape[1], first.states.shape[2], first.states.device
        times = np.zeros((B, T))
        states = torch.empty((B, T, d), dtype=torch.complex128, device=dev)
        probs = torch.empty((B, T, d), dtype=torch.float64, device=dev)
        pops = torch.empty((B, T, 2), dtype=torch.float64, device=dev)
        durs = np.zeros(B, dtype=np.int64)
        for dur, (sel, sol) in parts.items():
            idx = torch.as_tensor(sel, device=dev)
            times[sel] = sol.times
            states[idx], probs[idx], pops[idx] = sol.states, sol.probabilities, sol.populations
            durs[sel] = dur
        return DurationSweepSolution(times=times, states=states, probabilities=probs, populations=pops, durations=durs)


class B200PulseBackend(PulseBackend):
    """PulseBackend whose ``solve`` runs on a B200 through libcasq_b200.so (no CPU fallback)."""

    def __init__(self, hamiltonian: HamiltonianModel, control: ControlModel, seed: Optional[int] = None, device: int = 0,
                 noise=None, forward_subsystem_dims: bool = False):
        """``forward_subsystem_dims``: casq builds its DynamicsBackend WITHOUT subsystem_dims
        (src/casq/backends/qiskit/qiskit_pulse_backend.py:184-188), so upstream treats a multi-qubit model as ONE subsystem
        of dimension prod(dims) and ``solve`` folds every level >= 1 into "1" (SURVEY section 5 quirk 2).  False (default)
        keeps that behaviour bit for bit; True forwards the dimensions: populations / counts are per measured qubit
        (bit-strings, levels >= 1 of each qubit clipped to "1") and ``Solution.qubits`` lists the measured qubits."""
        self._device = int(device)
        self.noise = noise
        self.forward_subsystem_dims = bool(forward_subsystem_dims)
        super().__init__(hamiltonian=hamiltonian, control=control, seed=seed)
        if noise is not None:
            self._native_backend.set_dissipators(noise.lindblad_operators(self.hamiltonian.qubit_dims))

    # ------------------------------------------------------------------ construction
    def _get_native_backend(self) -> NativeModel:
        freqs = self.control.channel_carrier_freqs or {}
        missing = [c for c in self.hamiltonian.channels if c not in freqs]
        if missing:
            raise CasqError(f"ControlModel.channel_carrier_freqs has no entry for channel(s) {missing}.")
        return NativeModel(
            static=self.hamiltonian.static_operator,
            operators=self.hamiltonian.operators,
            carrier_hz=[freqs[c] for c in self.hamiltonian.channels],
            dt=self.control.dt,
            rotating_frame=self.hamiltonian.rotating_frame,
            rwa_cutoff_freq=self.hamiltonian.rwa_cutoff_freq,
            device=self._device,
        )

    @property
    def native(self) -> NativeModel:
        return self._native_backend

    @property
    def dim(self) -> int:
        return self._native_backend.dim

    def ground_state(self) -> np.ndarray:
        """Column 0 of the dressed eigenbasis: what ``initial_state=None`` means."""
        return np.array(self._native_backend.dressed_states[:, 0])

    # ------------------------------------------------------------------ schedule handling
    def _circuit_to_schedule(self, circuit: PulseCircuit):
        """ASAP-schedule the calibrated gates; returns (Schedule, t_acquire in samples, measured qubits)."""
        cursor: dict = {}
        sched = Schedule(name=circuit.name)
        acquire_times = []
        measured = []
        for inst in circuit.data:
            if inst.operation == "measure":
                q = inst.qubits[0]
                acquire_times.append(cursor.get(q, 0))
                measured.append(q)
                continue
            gate = inst.operation
            cal = circuit.calibrations.get(gate.name, {}).get((tuple(inst.qubits), ()))
            if cal is None:
                raise CasqError(f"Gate '{gate.name}' on qubits {inst.qubits} has no pulse calibration.")
            start = max(cursor.get(q, 0) for q in inst.qubits)
            for s, sub in cal.instructions:
                sched.insert(start + s, sub)
            for q in inst.qubits:
                cursor[q] = start + cal.duration
        if not acquire_times:
            raise CasqError("At least one measurement must be present in the circuit (DynamicsBackend requires an acquire).")
        if len(set(acquire_times)) > 1:
            t_acq = max(acquire_times)  # measurements are aligned to the latest one
        else:
            t_acq = acquire_times[0]
        if t_acq <= 0:
            raise CasqError("The circuit has no pulse before its measurement: nothing to simulate.")
        return sched, t_acq, measured

    def _slots_from_schedule(self, sched: Schedule):
        """Play instructions -> engine slots.  Phase / frequency instructions act on the LATER plays of their channel, as in
        qiskit-dynamics' InstructionToSignals: the accumulated phase joins the pulse angle, the
        accumulated frequency shift becomes the slot's freq_shift."""
        slots, params = [], []
        phase: dict = {}
        fshift: dict = {}
        for start, inst in sorted(sched.instructions, key=lambda si: si[0]):
            ch = getattr(inst, "channel", None)
            name = ch.name if ch is not None else None
            if isinstance(inst, ShiftPhase):
                phase[name] = phase.get(name, 0.0) + inst.phase
                continue
            if isinstance(inst, SetPhase):
                phase[name] = inst.phase
                continue
            if isinstance(inst, ShiftFrequency):
                fshift[name] = fshift.get(name, 0.0) + inst.frequency
                continue
            if isinstance(inst, SetFrequency):
                fshift[name] = inst.frequency - (self.control.channel_carrier_freqs or {}).get(name, 0.0)
                continue
            if not isinstance(inst, Play):
                continue
            if name not in self.hamiltonian.channels:
                raise CasqError(f"Schedule plays on channel '{name}' which is not in the Hamiltonian model {self.hamiltonian.channels}.")
            slots.append((inst.pulse.kind, self.hamiltonian.channels.index(name), start, inst.pulse.duration))
            ep = inst.pulse.engine_params()
            if phase.get(name):
                ep["angle"] = np.asarray(ep.get("angle", 0.0), dtype=float) + phase[name]
            if fshift.get(name):
                ep["freq_shift"] = fshift[name]
            params.append(ep)
        if len(slots) > 4:
            raise CasqError(f"At most 4 Play instructions per schedule are supported, got {len(slots)}.")
        return slots, params

    # ------------------------------------------------------------------ core batched run
    def _run(self, slots, params_dev, y0, t_span, method: str, t_eval, options: dict):
        model = self._native_backend
        nsteps = status = None
        if method in _FIXED:
            if "max_dt" not in options:
                raise CasqError(f"Method '{method}' is a fixed-step solver and needs run_options['max_dt'].")
            raw = engine.solve_sv_rk4(model, slots, params_dev, y0, t_span, options["max_dt"], t_eval)
        elif method == "jax_odeint" or method in _SCIPY:
            if method in _SCIPY and "rtol" not in options and "atol" not in options:
                _log.warning("ODE method '%s' runs on the device Dopri5 with rtol = atol = 1.4e-8; casq's scipy path would run at "
                             "scipy's defaults (rtol 1e-3, atol 1e-6). Pass run_options={'rtol':..., 'atol':...} to choose.", method)
            # scipy-named adaptive methods run on the device Dopri5 as well.  casq passes no tolerances, so the
            # reference runs them at scipy's rtol=1e-3 (SURVEY §3.1 "accuracy trap"); here an unset tolerance
            # means 1.4e-8 (the jax_odeint default), i.e. closer to the true solution than the reference.
            raw, nsteps, status = engine.solve_sv_dopri5(
                model, slots, params_dev, y0, t_span,
                rtol=options.get("rtol", 1.4e-8), atol=options.get("atol", 1.4e-8),
                hmax=options.get("hmax", options.get("max_step", 0.0)), mxstep=options.get("mxstep", 0),
                t_eval=t_eval, return_stats=True)
        elif method == "LSODA":
            raise CasqError("SCIPY_LSODA cannot integrate complex states (scipy limitation in the reference too).")
        else:
            raise CasqError(f"Unknown ODE solver method '{method}'.")
        return raw, nsteps, status

    def solve_slots_batch(self, slots: Sequence[tuple], slot_params: Sequence[dict], t_final: int, method: str = "RK4",
                          initial_state=None, steps: Optional[int] = None, run_options: Optional[dict] = None,
                          keep_raw: bool = False, params_device: Optional[torch.Tensor] = None) -> BatchSolution:
        """Lowest-level batched entry: explicit slots [(kind, channel_index, start, duration)] and per-slot
        parameter dicts (scalars or length-B arrays), t_span = [0, t_final*dt]."""
        options = dict(run_options or {})
        options.pop("method", None)
        dev = torch.device("cuda", self._device)
        if params_device is None:
            if not torch.cuda.is_available():
                raise CasqError("casq_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
            params_device = engine.pack_params(slot_params, device=dev)
        t_span = (0.0, float(t_final) * self.control.dt)
        t_eval = None if steps is None else np.linspace(t_span[0], t_span[1], int(steps))
        y0 = self.ground_state() if initial_state is None else np.asarray(getattr(initial_state, "data", initial_state))
        if y0.ndim == 2 and y0.shape == (self.dim, self.dim):
            raise CasqError("DensityMatrix initial states need the Lindblad path (solve_lindblad_batch).")
        raw, nsteps, status = self._run(slots, params_device, y0, t_span, method, t_eval, options)
        times = np.asarray(t_span) if t_eval is None else t_eval
        states, probs, pops = engine.postprocess_sv(self._native_backend, times, raw)
        return BatchSolution(times=times, states=states, probabilities=probs, populations=pops,
                             raw=raw if keep_raw else None, nsteps=nsteps, status=status)

    def solve_propagators_batch(self, slots: Sequence[tuple], slot_params: Sequence[dict], t_final: int, max_dt: float,
                                params_device: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Full propagators [B, d, d] (rotating frame, lab basis) by exponential-midpoint steps of size <= max_dt:
        the piecewise-constant path (upstream "scipy_expm"/"jax_expm"; not reachable through casq's method enum)."""
        if params_device is None:
            params_device = engine.pack_params(slot_params, device=torch.device("cuda", self._device))
        return engine.propagators_expm(self._native_backend, slots, params_device, (0.0, float(t_final) * self.control.dt), max_dt)

    def gate_fidelities_batch(self, slots: Sequence[tuple], slot_params: Sequence[dict], t_final: int, max_dt: float, target,
                              levels: int = 2, virtual_z: bool = False, params_device: Optional[torch.Tensor] = None):
        """Average gate fidelity, process fidelity and leakage of every pulse point against ``target`` (a unitary on the
        computational subspace), from the batched propagators in the model's rotating frame (casq_b200.fidelity)."""
        from ...fidelity import gate_fidelities

        U = self.solve_propagators_batch(slots, slot_params, t_final, max_dt, params_device)
        qd = self.hamiltonian.qubit_dims
        return gate_fidelities(U, target, [int(qd[q]) for q in sorted(qd)], levels, virtual_z)

    def solve_lindblad_batch(self, slots: Sequence[tuple], slot_params: Sequence[dict], t_final: int, max_dt: float,
                             initial_state=None, steps: Optional[int] = None, want_states: bool = False,
                             params_device: Optional[torch.Tensor] = None):
        """Density-matrix evolution with the backend's NoiseModel (fixed-step RK4).  Returns (times, rho [B,T,d,d] in
        the rotating frame or None, populations [B,T,d] of the dressed states, trace-normalised like casq's
        DensityMatrix branch, helpers.py:110-121).  Needs a diagonal rotating frame (evaluation_mode=None models)."""
        if self.noise is None:
            raise CasqError("solve_lindblad_batch needs a backend built with a NoiseModel (noise=...).")
        if params_device is None:
            params_device = engine.pack_params(slot_params, device=torch.device("cuda", self._device))
        t_span = (0.0, float(t_final) * self.control.dt)
        t_eval = None if steps is None else np.linspace(t_span[0], t_span[1], int(steps))
        rho0 = self.ground_state() if initial_state is None else np.asarray(getattr(initial_state, "data", initial_state))
        rho, pops = engine.solve_lindblad_rk4(self._native_backend, slots, params_device, rho0, t_span, max_dt, t_eval,
                                              want_rho=want_states, want_pops=True)
        pops = pops / pops.sum(dim=-1, keepdim=True)
        return (np.asarray(t_span) if t_eval is None else t_eval), rho, pops

    def solve_batch(self, gate: PulseGate, parameters: Union[dict, np.ndarray], method: "PulseBackend.ODESolverMethod",
                    qubit: int = 0, amplitude=None, angle=None, initial_state=None, steps: Optional[int] = None,
                    run_options: Optional[dict] = None, keep_raw: bool = False, detuning=None, duration=None,
                    channel: Optional[str] = None) -> BatchSolution:
        """Sweep / candidate batch for one gate: ``parameters`` is a dict of arrays (e.g. {"sigma": [B],
        "beta": [B]}) or a [B, P] array in ``gate.to_parameters_dict`` order; ``amplitude``/``angle`` may be
        arrays too (default: the gate's own).  Reduces to ``solve`` semantics for B = 1.

        ``detuning`` (Hz, scalar or [B]): per-point frequency shift of the drive, i.e. a ShiftFrequency in front of the
        Play (samples times exp(i 2 pi detuning dt n)) -- one launch for the whole sweep.
        ``duration`` (samples, scalar or [B] integers): per-point gate duration (PulseGate.duration, pulse_gate.py:69-72).
        The duration fixes t_span = [0, duration dt] (dynamics_backend_patch.py:180-184) and with it the solver's time
        grid, so the batch runs as one launch per distinct duration, each on exactly the grid the reference would use for
        that duration; the result is a ``DurationSweepSolution`` (per-point times)."""
        if not isinstance(parameters, dict):
            arr = np.atleast_2d(np.asarray(parameters, dtype=float))
            parameters = gate.to_parameters_dict(arr.T)
        amp = gate.amplitude if amplitude is None else amplitude
        ang = gate.angle if angle is None else angle
        name = channel if channel is not None else gate.channel_name(qubit)
        if name not in self.hamiltonian.channels:
            raise CasqError(f"Channel '{name}' is not in the Hamiltonian model {self.hamiltonian.channels}.")
        m = method.value if isinstance(method, PulseBackend.ODESolverMethod) else str(method)
        kw = {k: np.asarray(v, dtype=float) for k, v in parameters.items()}

        def build(dur, sel=None):
            pick = (lambda v: v) if sel is None else (lambda v: v[sel] if np.ndim(v) > 0 else v)
            pulse = type(gate.pulse({k: pick(v) for k, v in kw.items()}))(
                duration=int(dur), amp=pick(np.asarray(amp, dtype=float)),
                angle=0.0 if ang is None else pick(np.asarray(ang, dtype=float)),
                limit_amplitude=gate.limit_amplitude, name=gate.name, **{k: pick(v) for k, v in kw.items()})
            ep = pulse.engine_params()
            if detuning is not None:
                ep["freq_shift"] = pick(np.asarray(detuning, dtype=float))
            return pulse, ep

        if duration is None or np.ndim(duration) == 0:
            pulse, ep = build(gate.duration if duration is None else duration)
            slots = [(pulse.kind, self.hamiltonian.channels.index(name), 0, pulse.duration)]
            return self.solve_slots_batch(slots, [ep], pulse.duration, m, initial_state, steps, run_options, keep_raw)
        durations = np.asarray(duration)
        if not np.all(durations == np.round(durations)) or np.any(durations < 1):
            raise CasqError("Per-point durations must be positive integers (samples).")
        durations = durations.astype(np.int64)
        B = len(durations)
        parts = {}
        for dur in np.unique(durations):
            sel = np.nonzero(durations == dur)[0]
            pulse, ep = build(dur, sel)
            slots = [(pulse.kind, self.hamiltonian.channels.index(name), 0, pulse.duration)]
            y0 = initial_state
            if y0 is not None and np.ndim(getattr(y0, "data", y0)) == 2 and np.shape(getattr(y0, "data", y0))[0] == B:
                y0 = np.asarray(getattr(y0, "data", y0))[sel]
            parts[int(dur)] = (sel, self.solve_slots_batch(slots, [ep], int(dur), m, y0, steps, run_options, keep_raw))
        return DurationSweepSolution.assemble(B, parts)

    # ------------------------------------------------------------------ the drop-in call
    def solve(
        self,
        circuit: PulseCircuit,
        method: "PulseBackend.ODESolverMethod",
        initial_state: Optional[Union[DensityMatrix, Statevector]] = None,
        shots: int = 1024,
        steps: Optional[int] = None,
        run_options: Optional[dict[str, Any]] = None,
    ) -> "PulseBackend.Solution":
        """Same contract as QiskitPulseBackend.solve (qiskit_pulse_backend.py:119-164).

        The enum ``method`` wins over run_options["method"] (:147-150); ``initial_state=None`` is the dressed
        ground state; ``steps`` gives that many evenly spaced time points including both ends
        (dynamics_backend_patch.py:180-184), otherwise the two end points are returned.
        """
        if run_options:
            run_options.update(method=method.value)
        else:
            run_options = {"method": method.value}
        sched, t_acq, measured = self._circuit_to_schedule(circuit)
        slots, params = self._slots_from_schedule(sched)
        y0 = None if initial_state is None else np.asarray(getattr(initial_state, "data", initial_state))
        if isinstance(initial_state, DensityMatrix) or (y0 is not None and y0.ndim == 2):
            return self._solve_density_matrix(circuit.name, slots, params, t_acq, method.value, y0, shots, steps, run_options, measured)
        batch = self.solve_slots_batch(slots, params, t_acq, method.value, initial_state, steps, run_options)
        if batch.status is not None and int(batch.status[0].item()) != 0:
            raise CasqError(f"Adaptive solve failed for this circuit (status {int(batch.status[0].item())}).")
        states = batch.states[0].cpu().numpy()
        probs = batch.probabilities[0].cpu().numpy()
        return self._to_solution(circuit.name, batch.times, states, probs, shots, measured)

    def _solve_density_matrix(self, name, slots, params, t_acq, method, rho0, shots, steps, run_options, measured):
        """The DensityMatrix branch of get_experiment_result (src/casq/backends/qiskit/helpers.py:110-121): the Lindblad
        kernels integrate rho (with the backend's NoiseModel, or without dissipators), then per time point
        rho <- V^+ (e^{tF} rho e^{-tF}) V and the trace is normalised."""
        if method not in _FIXED:
            raise CasqError("A DensityMatrix initial state is integrated by the fixed-step Lindblad path: use "
                            "ODESolverMethod.QISKIT_DYNAMICS_RK4 with run_options['max_dt'].")
        options = dict(run_options or {})
        if "max_dt" not in options:
            raise CasqError(f"Method '{method}' is a fixed-step solver and needs run_options['max_dt'].")
        d = self.dim
        if rho0.shape != (d, d):
            raise CasqError(f"DensityMatrix initial state must be [{d},{d}], got {rho0.shape}.")
        model = self._native_backend
        t_span = (0.0, float(t_acq) * self.control.dt)
        t_eval = None if steps is None else np.linspace(t_span[0], t_span[1], int(steps))
        pdev = engine.pack_params(params, device=torch.device("cuda", self._device))
        rho, _ = engine.solve_lindblad_rk4(model, slots, pdev, rho0, t_span, options["max_dt"], t_eval, want_rho=True, want_pops=False)
        times = np.asarray(t_span) if t_eval is None else t_eval
        rho = rho[0].cpu().numpy()
        V, e = model.dressed_states, model.frame_evals  # the Lindblad path runs in a diagonal frame: frame basis = identity
        states, probs = [], []
        for k, t in enumerate(times):
            ph = np.exp(-1j * e * t)
            lab = ph[:, None] * rho[k] * ph.conj()[None, :]
            r = V.conj().T @ lab @ V
            r = r / np.trace(r).real
            states.append(r)
            probs.append(np.real(np.diag(r)))
        return self._to_solution(name, times, np.array(states), np.array(probs), shots, measured, density=True)

    def _to_solution(self, name, times, states, probs, shots, measured, density: bool = False) -> "PulseBackend.Solution":
        """get_experiment_result + convert_to_solution (helpers.py:58-230) for one experiment.

        Populations: helpers.py:129-134 (levels > 1 clipped to "1"); samples / counts: _sample_probability_dict and
        _get_counts_from_samples with the experiment seed (:137-141); IQ: _get_iq_data (:143-159) -- per measured
        subsystem a multinomial split of the shots over its levels, then PER LEVEL the I and the Q normal draws
        ([UPSTREAM] qiskit-dynamics 0.4.1 appends rng.normal for I and for Q inside the same level loop), centres on the
        unit circle at angle 2 pi level / dim, width 0.2, scaled by 2."""
        d = self.dim
        rng = np.random.default_rng(self._seed)
        seed = int(rng.integers(low=0, high=9223372036854775807))
        qdims = self.hamiltonian.qubit_dims
        labels = sorted(qdims)
        if self.forward_subsystem_dims and len(labels) > 1:
            dims = [int(qdims[q]) for q in labels]
            missing = [q for q in measured if q not in labels]
            if missing:
                raise CasqError(f"Measured qubit(s) {missing} are not in the Hamiltonian model {labels}.")
            meas_idx = [labels.index(q) for q in measured]
            qubits = list(measured)
        else:  # casq's behaviour: one subsystem of dimension prod(dims), reported as qubit 0
            dims, meas_idx, qubits = [d], [0], [0]
        out_states, pops_l, samples_l, counts_l, iq_l, avg_l = [], [], [], [], [], []
        State = DensityMatrix if density else Statevector
        pgrid = None
        for t in range(len(times)):
            out_states.append(State(states[t], dims=tuple(dims)))
            p = np.clip(probs[t], 0.0, None)
            pop = {k: v for k, v in populations_by_qubit(p, dims, meas_idx).items() if v > 0}
            pops_l.append(pop)
            keys = list(pop.keys())
            pv = np.array([pop[k] for k in keys])
            smp = np.random.default_rng(seed).choice(keys, size=shots, p=pv / pv.sum())
            samples_l.append(smp.tolist())
            vals, cnts = np.unique(smp, return_counts=True)
            counts_l.append({str(v): int(c) for v, c in zip(vals, cnts)})
            rng_iq = np.random.default_rng(seed)
            pgrid = p.reshape([dims[k] for k in reversed(range(len(dims)))])  # axis 0 = last subsystem
            full_i, full_q = [], []
            for sub in meas_idx:
                axis = len(dims) - 1 - sub
                psub = pgrid.sum(axis=tuple(a for a in range(len(dims)) if a != axis))
                theta = 2 * np.pi / dims[sub]
                n_lvl = rng_iq.multinomial(shots, psub / psub.sum())
                sub_i, sub_q = [], []
                for lvl, n in enumerate(n_lvl):
                    sub_i.append(rng_iq.normal(np.cos(lvl * theta), 0.2, n))
                    sub_q.append(rng_iq.normal(np.sin(lvl * theta), 0.2, n))
                full_i.append(np.concatenate(sub_i))
                full_q.append(np.concatenate(sub_q))
            iq = 2 * np.stack([np.array(full_i).T, np.array(full_q).T], axis=-1)  # [shots, slots, 2]
            if len(meas_idx) == 1:
                iq_l.append([(float(a), float(b)) for a, b in iq[:, 0, :]])
                avg = iq[:, 0, :].mean(axis=0)
                avg_l.append((float(avg[0]), float(avg[1])))
            else:
                iq_l.append(iq.tolist())
                avg_l.append(iq.mean(axis=0).tolist())
        return PulseBackend.Solution(
            circuit_name=name, qubits=qubits, times=np.asarray(times), samples=samples_l, counts=counts_l,
            populations=pops_l, states=out_states, iq_data=iq_l, avg_iq_data=avg_l, shots=shots, seed=seed,
            is_success=True, timestamp=datetime.timestamp(datetime.now()))
