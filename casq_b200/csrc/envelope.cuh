// casq_b200: on-device envelope sampler (device functions shared by the stand-alone sampler kernel and
// by every stepper, which evaluates the envelope on the fly when the sample index changes).
//
// Families = the qiskit.pulse.library classes casq's gates forward to (src/casq/gates/*_pulse_gate.py);
// formulas per SURVEY.md Appendix A.2: midpoint sampling samples[n] = envelope(n + 1/2),
// lifted(t; mu, t0, sigma) = (g(t) - g(t0)) / (1 - g(t0)), g(x) = exp(-((x-mu)/sigma)^2 / 2).
#pragma once
#include <cstdlib>

#include "common.cuh"

struct dcplx {
  double re, im;
};

__device__ __forceinline__ double casq_lifted(double t, double mu, double t_zero, double inv_sigma) {
  double a = (t - mu) * inv_sigma;
  double b = (t_zero - mu) * inv_sigma;
  double g = exp(-0.5 * a * a);
  double off = exp(-0.5 * b * b);
  return (g - off) / (1.0 - off);
}

// envelope(n + 1/2) for local sample n in [0, duration); params = (amp, angle, sigma, width, beta)
__device__ __forceinline__ dcplx casq_envelope_eval(int kind, int duration, int n, double amp, double angle,
                                                    double sigma, double width, double beta) {
  double sa, ca;
  sincos(angle, &sa, &ca);
  const double t = (double)n + 0.5;
  const double c = 0.5 * (double)duration;
  double er = 1.0, ei = 0.0;
  if (kind == CASQ_PULSE_GAUSSIAN || kind == CASQ_PULSE_DRAG) {
    const double inv_s = 1.0 / sigma;
    const double g = casq_lifted(t, c, (double)duration + 1.0, inv_s);
    er = g;
    if (kind == CASQ_PULSE_DRAG) ei = beta * (-(t - c) * inv_s * inv_s) * g;
  } else if (kind == CASQ_PULSE_GAUSSIAN_SQUARE || kind == CASQ_PULSE_GAUSSIAN_SQUARE_DRAG) {
    const double inv_s = 1.0 / sigma;
    const double tl = c - 0.5 * width, tr = c + 0.5 * width;
    if (t <= tl) {
      const double g = casq_lifted(t, tl, -1.0, inv_s);
      er = g;
      if (kind == CASQ_PULSE_GAUSSIAN_SQUARE_DRAG) ei = beta * (-(t - tl) * inv_s * inv_s) * g;
    } else if (t >= tr) {
      const double g = casq_lifted(t, tr, (double)duration + 1.0, inv_s);
      er = g;
      if (kind == CASQ_PULSE_GAUSSIAN_SQUARE_DRAG) ei = beta * (-(t - tr) * inv_s * inv_s) * g;
    }
  }
  dcplx out;
  const double ar = amp * ca, ai = amp * sa;
  out.re = ar * er - ai * ei;
  out.im = ar * ei + ai * er;
  return out;
}

// Per-point detuning: samples of a frequency-shifted channel carry exp(i 2 pi freq_shift dt n), n = global sample index
// (InstructionToSignals: times = dt * (start_sample + arange(n_samples))).
__device__ __forceinline__ dcplx casq_apply_freq_shift(dcplx e, double freq_shift, double dt, int gidx) {
  double sn, cs;
  sincos(6.283185307179586 * freq_shift * (dt * (double)gidx), &sn, &cs);
  dcplx o;
  o.re = e.re * cs - e.im * sn;
  o.im = e.re * sn + e.im * cs;
  return o;
}

// Slots as the kernels see them: `channel` already remapped to the ACTIVE channel index.
struct DevSlots {
  int n;
  double dt;  // sample period (the phase of a per-point frequency shift advances by 2 pi freq_shift dt per sample)
  int kind[CASQ_MAX_SLOTS], ach[CASQ_MAX_SLOTS], start[CASQ_MAX_SLOTS], duration[CASQ_MAX_SLOTS];
};

// Sum of all slots playing on each active channel at global sample `gidx` for batch point b.
template <int KA>
__device__ __forceinline__ void casq_channel_envelopes(const DevSlots& s, const double* __restrict__ params,
                                                       long long B, long long b, int gidx, dcplx* env) {
#pragma unroll
  for (int k = 0; k < KA; k++) env[k].re = env[k].im = 0.0;
  for (int i = 0; i < s.n; i++) {
    const int n = gidx - s.start[i];
    if (n < 0 || n >= s.duration[i]) continue;
    const double* p = params + (long long)i * CASQ_NPARAM * B + b;
    dcplx e = casq_envelope_eval(s.kind[i], s.duration[i], n, p[0], p[B], p[2 * B], p[3 * B], p[4 * B]);
    const double fs = p[5 * B];
    if (fs != 0.0) e = casq_apply_freq_shift(e, fs, s.dt, gidx);
#pragma unroll
    for (int k = 0; k < KA; k++)
      if (s.ach[i] == k) {
        env[k].re += e.re;
        env[k].im += e.im;
      }
  }
}

// Envelope samples of the trajectories [b0, b0 + Bc) of a batch of B, evaluated once per solve: out[k][idx][b - b0],
// idx in [0, n_total).  Used by the steppers whose lanes sit at DIFFERENT sample indices (adaptive Dopri5) or that
// need a new sample on every step (propagators with one step per sample): a coalesced 16 B look-up replaces 2 exp +
// sincos + divisions per call.  Callers walk large batches in windows (casq_env_table_window) to bound the table.
#define CASQ_ENV_TABLE_MAX_BYTES (1ull << 30)
static inline long long casq_env_table_window(int KA, int n_total, long long B) {
  unsigned long long cap = CASQ_ENV_TABLE_MAX_BYTES;
  if (const char* e = getenv("CASQ_ENV_TABLE_MAX_KB")) cap = 1024ull * strtoull(e, nullptr, 10);  // test knob: force several windows
  const unsigned long long per_point = (unsigned long long)KA * (unsigned long long)(n_total > 0 ? n_total : 1) * 16ull;
  const long long w = (long long)(cap / per_point);
  return w < 1 ? 1 : (w < B ? w : B);
}

template <int KA>
__global__ void casq_env_table_kernel(DevSlots slots, const double* __restrict__ params, long long B, long long b0, long long Bc,
                                      int n_total, double2* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Bc * n_total) return;
  const int idx = (int)(i / Bc);
  const long long bl = i - (long long)idx * Bc;
  dcplx env[KA];
  casq_channel_envelopes<KA>(slots, params, B, b0 + bl, idx, env);
#pragma unroll
  for (int k = 0; k < KA; k++) out[((long long)k * n_total + idx) * Bc + bl] = make_double2(env[k].re, env[k].im);
}

template <int KMAX>
static inline cudaError_t casq_launch_env_table(int KA, DevSlots slots, const double* params, long long B, long long b0, long long Bc,
                                        int n_total, double2* out, cudaStream_t st) {
  const long long n = Bc * n_total;
  if (n <= 0) return cudaSuccess;
  const unsigned g = (unsigned)((n + 255) / 256);
  switch (KA) {
    case 1: casq_env_table_kernel<1><<<g, 256, 0, st>>>(slots, params, B, b0, Bc, n_total, out); break;
    case 2: casq_env_table_kernel<(KMAX >= 2 ? 2 : 1)><<<g, 256, 0, st>>>(slots, params, B, b0, Bc, n_total, out); break;
    case 3: casq_env_table_kernel<(KMAX >= 3 ? 3 : 1)><<<g, 256, 0, st>>>(slots, params, B, b0, Bc, n_total, out); break;
    default: casq_env_table_kernel<(KMAX >= 4 ? 4 : 1)><<<g, 256, 0, st>>>(slots, params, B, b0, Bc, n_total, out); break;
  }
  return cudaGetLastError();
}
