// casq_b200: C-ABI entry points that validate arguments and dispatch to the sm_100a kernels.
#include <algorithm>
#include <cstring>

#include "common.cuh"
#include "envelope.cuh"

bool casq_small_supported(const casq_model* m, int n_active);
int casq_small_rk4(casq_model* m, int KA, const int* active, const DevSlots& slots, int n_total, long long B,
                   const double* params_dev, const double* y0_dev, int y0_batched, double t0, double t1, double max_dt,
                   int T, const double* t_eval, double* out_dev, cudaStream_t st);

bool casq_general_supported(const casq_model* m, int n_active);
int casq_general_rk4(casq_model* m, int KA, const int* active, const DevSlots& slots, int n_total, long long B,
                     const double* params_dev, const double* y0_dev, int y0_batched, double t0, double t1, double max_dt,
                     int T, const double* t_eval, double* out_dev, cudaStream_t st);

// ---------------------------------------------------------------------------------------------
// stand-alone envelope sampler
__global__ void envelope_sample_kernel(int kind, int duration, long long B, const double* __restrict__ params, int start,
                                       double dt, double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * duration) return;
  const long long b = i / duration;
  const int n = (int)(i - b * duration);
  dcplx e = casq_envelope_eval(kind, duration, n, params[b], params[B + b], params[2 * B + b], params[3 * B + b],
                               params[4 * B + b]);
  const double fs = params[5 * B + b];
  if (fs != 0.0) e = casq_apply_freq_shift(e, fs, dt, start + n);
  reinterpret_cast<double2*>(out)[i] = make_double2(e.re, e.im);
}

extern "C" int casq_envelope_sample(int kind, int duration, int64_t B, const double* params_dev, int start, double dt,
                                    double* out_dev, void* stream) {
  if (kind < 0 || kind > CASQ_PULSE_GAUSSIAN_SQUARE_DRAG) CASQ_FAIL(CASQ_ERR_INVALID, "unknown pulse kind %d", kind);
  if (duration < 1 || B < 0) CASQ_FAIL(CASQ_ERR_INVALID, "casq_envelope_sample: bad arguments");
  if (B == 0) return CASQ_OK;
  if (!params_dev || !out_dev) CASQ_FAIL(CASQ_ERR_INVALID, "casq_envelope_sample: NULL buffer");
  const long long n = (long long)B * duration;
  envelope_sample_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(kind, duration, B, params_dev,
                                                                                       start, dt, out_dev);
  CASQ_CUDA(cudaGetLastError());
  return CASQ_OK;
}

// ---------------------------------------------------------------------------------------------
// Validate slots and compress the channel list to the channels that actually play something
// (a channel with no instruction carries an all-zero signal upstream and is skipped here).
int casq_prepare_slots(const casq_model* m, int n_slots, const casq_pulse_slot* slots, DevSlots* ds, int* active,
                       int* n_active, int* n_total) {
  if (n_slots < 0 || n_slots > CASQ_MAX_SLOTS) CASQ_FAIL(CASQ_ERR_UNSUPPORTED, "n_slots %d outside [0,%d]", n_slots, CASQ_MAX_SLOTS);
  if (n_slots > 0 && !slots) CASQ_FAIL(CASQ_ERR_INVALID, "slots is NULL");
  memset(ds, 0, sizeof(*ds));
  ds->n = n_slots;
  ds->dt = m->dt;
  *n_active = 0;
  *n_total = 0;
  for (int i = 0; i < n_slots; i++) {
    const casq_pulse_slot& s = slots[i];
    if (s.kind < 0 || s.kind > CASQ_PULSE_GAUSSIAN_SQUARE_DRAG) CASQ_FAIL(CASQ_ERR_INVALID, "slot %d: unknown pulse kind %d", i, s.kind);
    if (s.channel < 0 || s.channel >= m->K) CASQ_FAIL(CASQ_ERR_INVALID, "slot %d: channel %d not in the model (K=%d)", i, s.channel, m->K);
    if (s.duration < 1 || s.start < 0) CASQ_FAIL(CASQ_ERR_INVALID, "slot %d: bad start/duration", i);
    int a = -1;
    for (int k = 0; k < *n_active; k++)
      if (active[k] == s.channel) a = k;
    if (a < 0) {
      a = (*n_active)++;
      active[a] = s.channel;
    }
    ds->kind[i] = s.kind;
    ds->ach[i] = a;
    ds->start[i] = s.start;
    ds->duration[i] = s.duration;
    *n_total = std::max(*n_total, s.start + s.duration);
  }
  return CASQ_OK;
}

extern "C" int casq_solve_sv_rk4(casq_model* m, int n_slots, const casq_pulse_slot* slots, int64_t B,
                                 const double* params_dev, const double* y0_dev, int y0_batched, double t0, double t1,
                                 double max_dt, int T, const double* t_eval, double* out_dev, void* stream) {
  if (!m) CASQ_FAIL(CASQ_ERR_INVALID, "casq_solve_sv_rk4: NULL model");
  if (B < 0) CASQ_FAIL(CASQ_ERR_INVALID, "casq_solve_sv_rk4: negative batch size");
  if (B > 0 && (!y0_dev || !out_dev || (n_slots > 0 && !params_dev))) CASQ_FAIL(CASQ_ERR_INVALID, "casq_solve_sv_rk4: NULL buffer");
  if (!(t1 > t0)) CASQ_FAIL(CASQ_ERR_INVALID, "casq_solve_sv_rk4: need t1 > t0");
  DevSlots ds;
  int active[CASQ_MAX_SLOTS], n_active = 0, n_total = 0;
  int rc = casq_prepare_slots(m, n_slots, slots, &ds, active, &n_active, &n_total);
  if (rc != CASQ_OK) return rc;
  if (B == 0) return CASQ_OK;
  CASQ_CUDA(cudaSetDevice(m->device));
  if (n_active == 0) {  // free evolution still needs one (silent) channel slot on the small path
    if (m->K < 1) CASQ_FAIL(CASQ_ERR_UNSUPPORTED, "model without channels");
    active[0] = 0;
    n_active = 1;
  }
  if (casq_small_supported(m, n_active))
    return casq_small_rk4(m, n_active, active, ds, n_total, B, params_dev, y0_dev, y0_batched, t0, t1, max_dt, T, t_eval,
                          out_dev, (cudaStream_t)stream);
  if (casq_general_supported(m, n_active))
    return casq_general_rk4(m, n_active, active, ds, n_total, B, params_dev, y0_dev, y0_batched, t0, t1, max_dt, T, t_eval,
                            out_dev, (cudaStream_t)stream);
  CASQ_FAIL(CASQ_ERR_UNSUPPORTED, "casq_solve_sv_rk4: dim %d with %d active channels is outside the built kernels (dim <= 256, <= 4 channels)", m->d, n_active);
}

// ---------------------------------------------------------------------------------------------
// fp64 FMA probe: 16 independent chains per thread, `iters` rounds.
__global__ void fp64_fma_probe_kernel(int iters, double* sink) {
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  const double m = 1.0 - 1e-12, c = 1e-13;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = fma(a[i], m, c);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += a[i];
  if (s == 123.456) sink[0] = s;  // never true; keeps the chains alive
}

extern "C" int casq_fp64_fma_probe(int grid, int block, int iters, double* sink_dev, void* stream) {
  if (grid < 1 || block < 1 || block > 1024 || iters < 1 || !sink_dev) CASQ_FAIL(CASQ_ERR_INVALID, "casq_fp64_fma_probe: bad arguments");
  fp64_fma_probe_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(iters, sink_dev);
  CASQ_CUDA(cudaGetLastError());
  return CASQ_OK;
}
