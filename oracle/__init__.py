"""CPU oracle for the casq pulse-simulation hot path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED.  The reference (sinel/casq, /root/reference) does no arithmetic of its
own on this path: every floating-point operation happens in un-vendored third-party
packages that are absent from /root/reference and not installable offline:

    qiskit-dynamics 0.4.1   (poetry.lock:3276-3277)   Solver / solve_lmde / RK4_solver / jax_odeint,
                                                      InstructionToSignals, DiscreteSignal, RotatingFrame,
                                                      rotating_wave_approximation, DynamicsBackend
    qiskit-terra    0.24.2  (poetry.lock:3439-3440)   pulse.library symbolic pulses, Statevector
    jax             0.4.6   (poetry.lock:1240-1241)   jax.experimental.ode.odeint (Dopri5)
    scipy           1.9.3   (poetry.lock:3931-3932)   solve_ivp, expm

and the reference's own tests pin no numerical result for the path (they only assert
``isinstance(solution, PulseBackend.Solution)``: tests/casq/backends/qiskit/
qiskit_pulse_backend_test.py:64,76).  This package restates the *published algorithms* of
those pinned versions in numpy/scipy, anchored on casq's call sites (each function cites the
casq file:line it follows), and is validated by library-independent known answers
(tests/test_oracle_known_answers.py): closed-form Rabi rotation, expm of a constant
generator, norm conservation, RK4 h^4 convergence, cross-integrator agreement.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this package, and only as the checker / the timed CPU arm.
The product (``casq_b200``) never imports it and has no CPU fallback.
"""
from .model import (  # noqa: F401
    OracleModel,
    parse_hamiltonian_dict,
    dressed_decomposition,
    MANILA,
    MANILA_CARRIERS,
    MANILA_DT,
    transmon_hamiltonian_dict,
)
from .pulses import (  # noqa: F401
    envelope_samples,
    floor_divide_index,
    PulseSpec,
    channel_samples,
)
from .solvers import (  # noqa: F401
    fixed_step_sizes,
    rk4_solve,
    dop853_solve,
    dopri5_jax_solve,
    expm_midpoint_propagators,
    lindblad_rk4_solve,
    lindblad_rhs_dense,
)
from .postprocess import (  # noqa: F401
    postprocess_statevectors,
    postprocess_density_matrices,
    populations_dict,
    sample_counts,
    casq_solve,
)
