/* casq_b200 -- C ABI of the B200-native pulse-simulation engine behind casq's PulseBackend.solve.
 *
 * The reference (sinel/casq) has NO FFI: its plugin boundary is the Python ABC
 *   PulseBackend.solve / _get_native_backend      src/casq/backends/pulse_backend.py:332-366
 * dispatched by build(BackendLibrary, ...)          src/casq/backends/helpers.py:47-72
 * and implemented once, on top of qiskit-dynamics, in
 *   QiskitPulseBackend._get_native_backend / solve  src/casq/backends/qiskit/qiskit_pulse_backend.py:119-189.
 * Every entry point below names the reference interface it replaces.  INTEGRATION.md shows the
 * ctypes binding a casq maintainer would add.
 *
 * Conventions
 *   - complex128 = interleaved (re, im) doubles; matrices row-major; lab frame, bare basis unless said.
 *   - *_dev pointers are CUDA device pointers on the model's device; everything else is host memory.
 *   - every launch goes on the caller's `stream` (a cudaStream_t passed as void*); the library never
 *     synchronises the device except where documented (host outputs).
 *   - every function returns 0 on success, a negative casq_status otherwise; casq_last_error() gives the
 *     message for the calling thread.  No exceptions, no Python, no torch types cross this ABI.
 *   - a model handle must not be used from two threads at once (the reference backend is not re-entrant
 *     either: solve mutates self._native_backend.steps, qiskit_pulse_backend.py:157).
 */
#ifndef CASQ_B200_H
#define CASQ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  CASQ_OK = 0,
  CASQ_ERR_INVALID = -1,     /* bad argument (what casq reports as CasqError / ValueError)          */
  CASQ_ERR_UNSUPPORTED = -2, /* shape outside what the kernels are built for                         */
  CASQ_ERR_CUDA = -3,        /* CUDA runtime failure (message carries cudaGetErrorString)            */
  CASQ_ERR_NOMEM = -4
} casq_status;

/* Envelope families = the qiskit.pulse.library classes the five casq gates forward to
 * (src/casq/gates/constant_pulse_gate.py:71, gaussian_pulse_gate.py:73, drag_pulse_gate.py:75,
 *  gaussian_square_pulse_gate.py:76, gaussian_square_drag_pulse_gate.py:76). */
typedef enum {
  CASQ_PULSE_CONSTANT = 0,
  CASQ_PULSE_GAUSSIAN = 1,
  CASQ_PULSE_DRAG = 2,
  CASQ_PULSE_GAUSSIAN_SQUARE = 3,
  CASQ_PULSE_GAUSSIAN_SQUARE_DRAG = 4
} casq_pulse_kind;

/* Per-point pulse parameters are a structure-of-arrays block of CASQ_NPARAM rows per slot:
 *   params_dev[(slot*CASQ_NPARAM + j)*B + b],  j = 0 amp, 1 angle, 2 sigma, 3 width, 4 beta, 5 freq_shift (Hz).
 * freq_shift is the per-point detuning of the Play instruction: qiskit-dynamics' InstructionToSignals multiplies the
 * samples of a channel whose frequency was shifted (ShiftFrequency / SetFrequency before the Play) by
 * exp(i 2 pi freq_shift dt n), n = the GLOBAL sample index (the code casq calls at src/casq/common/helpers.py:124-132);
 * 0 = no shift.  Per-point DURATION is not a row: the duration fixes t_span (PulseGate.duration,
 * src/casq/gates/pulse_gate.py:69-72 -> dynamics_backend_patch.py:180-184) and with it the fixed-step grid, so points of
 * different duration run as one launch per distinct duration (casq_b200.backends.b200: solve_batch(duration=[...])). */
#define CASQ_NPARAM 6
#define CASQ_MAX_SLOTS 4

/* One Play instruction of the schedule (PulseGate.schedule, src/casq/gates/pulse_gate.py:97-112):
 * which family, on which Hamiltonian channel (index into the model's channel list), from which sample. */
typedef struct {
  int32_t kind;     /* casq_pulse_kind */
  int32_t channel;  /* index into the model's channels */
  int32_t start;    /* first sample */
  int32_t duration; /* samples */
} casq_pulse_slot;

typedef struct casq_model casq_model;

const char* casq_last_error(void);
int casq_version(void);

/* ---- model ------------------------------------------------------------------------------------
 * Replaces QiskitPulseBackend._get_native_backend (qiskit_pulse_backend.py:166-189): the
 * Solver(static_hamiltonian, hamiltonian_operators, hamiltonian_channels, channel_carrier_freqs, dt,
 * rotating_frame, rwa_cutoff_freq) construction plus DynamicsBackend's dressed-state decomposition.
 *   static_h   [dim*dim] c128, rad/s          operators [n_channels*dim*dim] c128
 *   carrier_hz [n_channels] Hz                 dt seconds
 *   frame_kind 0 = lab frame, 1 = Hermitian matrix frame [dim*dim] c128, 2 = real diagonal [dim]
 *   rwa_cutoff_hz: NaN = no rotating-wave approximation
 * Host-only work (eigendecompositions, RWA masks); device buffers are created lazily by the first solve. */
int casq_model_create(casq_model** out, int dim, int n_channels, const double* static_h,
                      const double* operators, const double* carrier_hz, double dt, int frame_kind,
                      const double* frame, double rwa_cutoff_hz, int device);
void casq_model_destroy(casq_model* m);

/* Static Lindblad operators L_j (lab frame) [n*dim*dim] c128 -- the NoiseModel casq's object model
 * names but never builds (docs/source/_static/images/casq-object-model.gv:9). */
int casq_model_set_dissipators(casq_model* m, int n, const double* ops);

/* Read back what the model derived (all host buffers):
 *   frame_evals [dim], frame_basis [dim*dim] c128 (columns = eigenvectors),
 *   dressed_evals [dim], dressed_states [dim*dim] c128 (column p = dressed state with max overlap on bare p;
 *   phase fixed so that its largest component is real positive).  Any pointer may be NULL. */
int casq_model_get(const casq_model* m, double* frame_evals, double* frame_basis, double* dressed_evals,
                   double* dressed_states);
int casq_model_dim(const casq_model* m);
/* Drop the cached stage table (carrier / Bohr-phase table of the last fixed-step time grid) so that the next
 * solve rebuilds it; benchmarks call this every step so no precomputation is amortised outside the timing. */
int casq_model_invalidate_cache(casq_model* m);

/* ---- envelope sampler -------------------------------------------------------------------------
 * Replaces the per-solve InstructionToSignals.get_signals sampling (same code casq calls at
 * src/casq/common/helpers.py:124-127): out_dev[b*duration + n] = envelope(n + 1/2; params of point b). */
int casq_envelope_sample(int kind, int duration, int64_t B, const double* params_dev /*[CASQ_NPARAM*B] SoA*/,
                         int start /* first global sample of the Play (phase of a frequency shift) */, double dt,
                         double* out_dev /*[B*duration] c128*/, void* stream);

/* ---- state-vector solves ----------------------------------------------------------------------
 * Replace the Solver.solve call behind QiskitPulseBackend.solve (qiskit_pulse_backend.py:158-162) for a
 * batch of B pulse-parameter points sharing one schedule structure.
 *   y0_dev   [dim] (y0_batched=0) or [B*dim] c128, lab basis, state at t0 (frame transform is identity at t0=0)
 *   t_eval   host [T] or NULL (-> outputs at t0 and t1, T must be 2)
 *   out_dev  [B*T*dim] c128: states "in the rotating frame, lab basis" = what Solver.solve returns.
 * Both solvers: dim <= 4 with <= 2 driven channels on the register kernels, otherwise dim <= 256 with <= 4 driven
 * channels (sparse operator rows in shared memory: above dim ~64 that needs a diagonal rotating frame, casq's
 * from_backend default; CASQ_ERR_UNSUPPORTED otherwise).
 * casq_solve_sv_rk4: ODESolverMethod.QISKIT_DYNAMICS_RK4 (pulse_backend.py:63): classical RK4, the span
 * cut at every t_eval point, each piece in n = ceil-ish(|D|/max_dt) equal steps (upstream rule). */
int casq_solve_sv_rk4(casq_model* m, int n_slots, const casq_pulse_slot* slots, int64_t B,
                      const double* params_dev, const double* y0_dev, int y0_batched, double t0, double t1,
                      double max_dt, int T, const double* t_eval, double* out_dev, void* stream);

/* casq_solve_sv_dopri5: ODESolverMethod.QISKIT_DYNAMICS_JAX_ODEINT (pulse_backend.py:65): adaptive
 * Dormand-Prince 5(4) with jax.experimental.ode.odeint's controller (Hairer initial step, FSAL, RMS error
 * ratio, factor clip [0.2, 10], safety 0.9, dense output by 4th-order interpolant).  hmax <= 0 -> inf,
 * mxstep <= 0 -> unbounded (per output interval, like jax).  Optional per-point outputs:
 *   nsteps_dev [B*2] int64 (attempted, accepted), status_dev [B] int32 (0 ok, 1 mxstep hit, 2 step underflow;
 *   the remaining outputs of a failed point are NaN). */
int casq_solve_sv_dopri5(casq_model* m, int n_slots, const casq_pulse_slot* slots, int64_t B,
                         const double* params_dev, const double* y0_dev, int y0_batched, double t0, double t1,
                         double rtol, double atol, double hmax, int64_t mxstep, int T, const double* t_eval,
                         double* out_dev, int64_t* nsteps_dev, int32_t* status_dev, void* stream);

/* ---- piecewise-constant propagators ---------------------------------------------------------------
 * casq_propagators_expm: U = prod_n expm(h G_F(t_n + h/2)), n over the fixed-step grid of [t0, t1] with
 * h <= max_dt (same step-count rule as RK4): the exponential-midpoint scheme of qiskit-dynamics'
 * "scipy_expm"/"jax_expm" solvers.  NOT selectable through casq's ODESolverMethod enum
 * (src/casq/backends/pulse_backend.py:60-71) -- additive API for full-unitary sweeps (BASELINE cfg 3).
 * expm = scaling-and-squaring Pade (Higham 2005; orders 3/5/7/9, scaled to theta_9 above it).  dim <= 64.
 *   out_dev [B*dim*dim] c128: propagators in the rotating frame, lab basis (apply to the state at t0). */
int casq_propagators_expm(casq_model* m, int n_slots, const casq_pulse_slot* slots, int64_t B,
                          const double* params_dev, double t0, double t1, double max_dt, double* out_dev,
                          void* stream);

/* ---- Lindblad master equation --------------------------------------------------------------------
 * casq_solve_lindblad_rk4: fixed-step RK4 (same stepping rule as casq_solve_sv_rk4) of
 *   rho' = -i[H(t), rho] + sum_j (L_j rho L_j^+ - 1/2 {L_j^+ L_j, rho})
 * in the model's rotating frame with the dissipators of casq_model_set_dissipators -- qiskit-dynamics'
 * LindbladModel semantics; casq itself never passes dissipators (qiskit_pulse_backend.py:174-183), so this is
 * additive API (BASELINE cfg 4).  Needs a diagonal rotating frame (frame_kind 0 or 2): the superoperator is
 * applied in structured form in the bare tensor-product basis.
 *   rho0_dev     [dim*dim] or [B*dim*dim] c128 (lab frame = frame at t0 = 0)
 *   out_rho_dev  [B*T*dim*dim] c128, in the rotating frame (what Solver.solve would return), or NULL
 *   out_pops_dev [B*T*dim] f64: populations of the dressed states after frame removal, NOT trace-normalised
 *                (diag of V^+ e^{tF} rho e^{-tF} V, src/casq/backends/qiskit/helpers.py:110-121), or NULL */
int casq_solve_lindblad_rk4(casq_model* m, int n_slots, const casq_pulse_slot* slots, int64_t B,
                            const double* params_dev, const double* rho0_dev, int rho0_batched, double t0, double t1,
                            double max_dt, int T, const double* t_eval, double* out_rho_dev, double* out_pops_dev,
                            void* stream);

/* ---- result post-processing ---------------------------------------------------------------------
 * Replaces the per-time-point loop of get_experiment_result (src/casq/backends/qiskit/helpers.py:98-135)
 * for state vectors: raw_dev [B*T*dim] c128 as returned by the solves above, times host [T].
 *   states_dev [B*T*dim] c128  lab frame, dressed basis, normalised when `normalize`   (may be NULL)
 *   probs_dev  [B*T*dim] f64   level probabilities                                      (may be NULL)
 *   pops_dev   [B*T*2]   f64   casq populations {"0": level 0, "1": every other level}  (may be NULL) */
int casq_postprocess_sv(casq_model* m, int64_t B, int T, const double* times, const double* raw_dev,
                        int normalize, double* states_dev, double* probs_dev, double* pops_dev, void* stream);

/* ---- measurement support ----------------------------------------------------------------------
 * casq_fp64_fma_probe: a pure DFMA kernel used by bench.py to MEASURE the fp64 roofline denominator
 * (MEASURED_PEAKS.json has none).  Launches grid*block threads doing iters*16 dependent-chain FMAs;
 * flops = 2*16*iters*grid*block. */
int casq_fp64_fma_probe(int grid, int block, int iters, double* sink_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CASQ_B200_H */
