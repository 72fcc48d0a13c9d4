// casq_b200: adaptive Dormand-Prince 5(4) state-vector stepper for small Hilbert spaces (dim 2..4), one
// trajectory per thread, per-thread step control.
//
// Replaces the "jax_odeint" method behind QiskitPulseBackend.solve (ODESolverMethod.QISKIT_DYNAMICS_JAX_ODEINT,
// src/casq/backends/pulse_backend.py:65; the optimizer's documented method, docs/source/user_guide/
// pulse_optimizer.rst:44 with {"atol":1e-6,"rtol":1e-8,"hmax":dt}).  The control flow follows the PUBLISHED
// algorithm of jax.experimental.ode.odeint (jax 0.4.6, not in /root/reference):
// Hairer initial step (order 4), FSAL, error ratio = RMS(err/(atol+rtol*max(|y0|,|y1|))), accept iff <= 1,
// dt <- dt*min(10, max(0.9*ratio^-1/5, ratio<1 ? 1 : 0.2)) clipped to [0,hmax], steps run PAST each output
// time and outputs come from the 4th-order interpolant of the last accepted step.
//
// Unlike the fixed-step kernel the time points differ per trajectory, so carrier and Bohr phases are
// evaluated with fp64 sincos per stage (phases reach ~2e3 rad: fp32 is not an option).
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "envelope.cuh"
#include "dopri5_common.cuh"

#define DP_MAXD 4
#define DP_MAXK 2
#define DP_MAXE 6

struct Dopri5Args {
  long long B;
  int T;                 // merged time points (t0 ... t1)
  const double* tpts;    // [T] device
  const char* keep;      // [T]
  int T_out;             // number of kept points
  const double* params;  // SoA
  const double* y0;
  int y0_batched;
  double* out;        // [B][T_out][D] c128
  long long* nsteps;  // [B][2] attempted, accepted (may be NULL)
  int* status;        // [B] 0 ok, 1 mxstep hit, 2 dt underflow (may be NULL)
  const double2* env_tab;  // [KA][n_total][B] envelope samples of every trajectory (casq_env_table_kernel)
  int n_total;
  double dt_sample, inv_dt_sample;
  double rtol, atol, hmax;
  long long mxstep;
  int rwa;
  unsigned emask;  // which upper entries carry anything
  int has_diag;
  double two_pi_nu[DP_MAXK];
  double bohr[DP_MAXE];
  double freq_max;                       // max |phase_freq|: |t| freq_max < DP_TRIG_MAX selects the branch-free sincos
  double phase_freq[DP_MAXK + DP_MAXE];  // the phases one stage needs, in order: carriers, then the Bohr frequency of every
                                         // element the kernel variant carries (ladder elements only for TRI; 0 where masked)
  double Sr[DP_MAXE], Si[DP_MAXE];
  double Pr[DP_MAXK][DP_MAXE], Pi[DP_MAXK][DP_MAXE];
  double Nr[DP_MAXK][DP_MAXE], Ni[DP_MAXK][DP_MAXE];
  double Sd[DP_MAXD];
  double Pdr[DP_MAXK][DP_MAXD], Pdi[DP_MAXK][DP_MAXD];
  int identityU;
  double Ur[DP_MAXD][DP_MAXD], Ui[DP_MAXD][DP_MAXD];
};

// Envelope samples come from a per-solve table (casq_env_table_kernel, envelope.cuh): with per-thread stage times
// almost every RHS call of a warp has SOME lane crossing a sample boundary, so the on-the-fly evaluation (2 exp,
// sincos, divisions) ran on every call.  Neighbouring lanes sit in the same or adjacent samples: coalesced look-ups.

// Everything of the right-hand side that depends on t only: h_d[i] (real diagonal) and the upper-triangle elements
// h[e] = (S + sum_k z_k P_k + conj(z_k) N_k)[i,j] e^{i(e_i - e_j)t} of the frame-rotated Hamiltonian.
// TRI: only the ladder elements (i, i+1) exist (a transmon drive a + a^dagger in its own eigenbasis) -- D-1 Bohr phases
// instead of D(D-1)/2, and the apply skips the rest statically.
template <int D>
struct DpCoef {
  double hd[D];
  double hr[D * (D - 1) / 2], hi[D * (D - 1) / 2];
};

template <int D>
__host__ __device__ constexpr int dp_tri_e(int i) {  // index of (i, i+1) in the row-major upper-triangle enumeration
  return i * (2 * D - i - 1) / 2;
}
template <int D, int KA, bool TRI>
struct DpShape {
  static constexpr int E = D * (D - 1) / 2;
  static constexpr int EP = TRI ? D - 1 : E;  // elements carried
  static constexpr int NPH = KA + EP;         // phases per stage time
};

// the library routine (any argument), kept out of line: 15 inlined copies of its slow path made the step's code 3x larger
static __device__ __noinline__ double2 dp_sincos_slow(double x) {
  double s, c;
  sincos(x, &s, &c);
  return make_double2(s, c);
}

// sin / cos of N phases; `fast`: every argument is inside the range of the branch-free routine (the caller tests
// |t| max|freq| once instead of reducing over the arguments)
template <int N>
__device__ __forceinline__ void dp_sincos_n(bool fast, const double* x, double* sn, double* cs) {
  if (fast) {
#pragma unroll
    for (int n = 0; n < N; n++) dp_sincos_fast(x[n], &sn[n], &cs[n]);
  } else {
#pragma unroll
    for (int n = 0; n < N; n++) {
      const double2 r = dp_sincos_slow(x[n]);
      sn[n] = r.x;
      cs[n] = r.y;
    }
  }
}
__device__ __forceinline__ bool dp_fast_range(const Dopri5Args& a, double t) { return fabs(t) * a.freq_max < DP_TRIG_MAX; }

// envelope samples of the driven channels at time t (zero outside the schedule)
template <int KA>
__device__ __forceinline__ void dp_env(const Dopri5Args& a, long long b, double t, double2* env) {
  const int gidx = dev_sample_index_fast(t, a.dt_sample, a.inv_dt_sample, a.n_total);
  const bool inside = gidx >= 0 && gidx < a.n_total;
#pragma unroll
  for (int k = 0; k < KA; k++)
    env[k] = inside ? __ldg(a.env_tab + ((long long)k * a.n_total + gidx) * a.B + b) : make_double2(0.0, 0.0);
}

// The coefficients at one time from its envelope samples and the NPH sines / cosines of a.phase_freq[n] * t.
template <int D, int KA, bool TRI>
__device__ __forceinline__ void dp_build(const Dopri5Args& a, const double2* env, const double* sn, const double* cs,
                                         DpCoef<D>& cf) {
  constexpr int EP = DpShape<D, KA, TRI>::EP;
  double zr[KA], zi[KA];
#pragma unroll
  for (int k = 0; k < KA; k++) {
    zr[k] = env[k].x * cs[k] - env[k].y * sn[k];
    zi[k] = env[k].x * sn[k] + env[k].y * cs[k];
  }
#pragma unroll
  for (int i = 0; i < D; i++) {
    double h = 0.0;
    if (a.has_diag) {
      h = a.Sd[i];
#pragma unroll
      for (int k = 0; k < KA; k++) h += 2.0 * (zr[k] * a.Pdr[k][i] - zi[k] * a.Pdi[k][i]);
    }
    cf.hd[i] = h;
  }
#pragma unroll
  for (int p = 0; p < EP; p++) {
    const int e = TRI ? dp_tri_e<D>(p) : p;
    double gr = a.Sr[e], gi = a.Si[e];
#pragma unroll
    for (int k = 0; k < KA; k++) {
      // z*P + conj(z)*N
      gr += zr[k] * (a.Pr[k][e] + a.Nr[k][e]) - zi[k] * (a.Pi[k][e] - a.Ni[k][e]);
      gi += zr[k] * (a.Pi[k][e] + a.Ni[k][e]) + zi[k] * (a.Pr[k][e] - a.Nr[k][e]);
    }
    cf.hr[e] = gr * cs[KA + p] - gi * sn[KA + p];
    cf.hi[e] = gr * sn[KA + p] + gi * cs[KA + p];
  }
}

template <int D, int KA, bool TRI>
__device__ __forceinline__ void dp_coef(const Dopri5Args& a, const double2* env, double t, DpCoef<D>& cf) {
  constexpr int NPH = DpShape<D, KA, TRI>::NPH;
  double x[NPH], sn[NPH], cs[NPH];
#pragma unroll
  for (int n = 0; n < NPH; n++) x[n] = a.phase_freq[n] * t;
  dp_sincos_n<NPH>(dp_fast_range(a, t), x, sn, cs);
  dp_build<D, KA, TRI>(a, env, sn, cs, cf);
}

// k = -i H y
template <int D, bool TRI>
__device__ __forceinline__ void dp_apply(unsigned emask, const DpCoef<D>& cf, const double* yr, const double* yi, double* kr,
                                         double* ki) {
  double wr[D], wi[D];
#pragma unroll
  for (int i = 0; i < D; i++) {
    wr[i] = cf.hd[i] * yr[i];
    wi[i] = cf.hd[i] * yi[i];
  }
  int e = 0;
#pragma unroll
  for (int i = 0; i < D; i++) {
#pragma unroll
    for (int j = i + 1; j < D; j++) {
      if (TRI ? (j == i + 1) : ((emask & (1u << e)) != 0)) {
        const double hr = cf.hr[e], hi = cf.hi[e];
        wr[i] += hr * yr[j] - hi * yi[j];
        wi[i] += hr * yi[j] + hi * yr[j];
        wr[j] += hr * yr[i] + hi * yi[i];
        wi[j] += hr * yi[i] - hi * yr[i];
      }
      e++;
    }
  }
#pragma unroll
  for (int i = 0; i < D; i++) {
    kr[i] = wi[i];
    ki[i] = -wr[i];
  }
}

template <int D, int KA, bool TRI>
__device__ __forceinline__ void dp_rhs(const Dopri5Args& a, long long b, double t, const double* yr, const double* yi,
                                       double* kr, double* ki) {
  DpCoef<D> cf;
  double2 env[KA];
  dp_env<KA>(a, b, t, env);
  dp_coef<D, KA, TRI>(a, env, t, cf);
  dp_apply<D, TRI>(a.emask, cf, yr, yi, kr, ki);
}

__constant__ double DP_ALPHA[6] = {1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1.0, 1.0};
__constant__ double DP_BETA[6][6] = {
    {1.0 / 5, 0, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656, 0},
    {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84}};
__constant__ double DP_CERR[7] = {35.0 / 384 - 1951.0 / 21600, 0, 500.0 / 1113 - 22642.0 / 50085, 125.0 / 192 - 451.0 / 720,
                                  -2187.0 / 6784 - -12231.0 / 42400, 11.0 / 84 - 649.0 / 6300, -1.0 / 60};
__constant__ double DP_CMID[7] = {6025192743.0 / 30085553152.0 / 2, 0, 51252292925.0 / 65400821598.0 / 2,
                                  -2691868925.0 / 45128329728.0 / 2, 187940372067.0 / 1594534317056.0 / 2,
                                  -1776094331.0 / 19743644256.0 / 2, 11237099.0 / 235043384.0 / 2};

// One thread per trajectory.  (Two or four lanes per trajectory sharing the sin/cos work were tried for the optimizer's
// batch sizes, where only half of the schedulers have a warp: bit-identical results, but the redundant remainder of the
// step cost more than the split saved -- 3.97 / 4.68 ms against 2.95 ms at B = 1e4.)
template <int D, int KA, bool TRI>
__global__ void __launch_bounds__(64) sv_dopri5_kernel(const Dopri5Args a) {
  constexpr int NPH = DpShape<D, KA, TRI>::NPH;
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  double yr[D], yi[D];
  {
    const double* src = a.y0 + (a.y0_batched ? b * 2 * D : 0);
#pragma unroll
    for (int i = 0; i < D; i++) {
      double sr = 0, si = 0;
#pragma unroll
      for (int j = 0; j < D; j++) {
        const double ur = a.identityU ? (i == j ? 1.0 : 0.0) : a.Ur[j][i];
        const double ui = a.identityU ? 0.0 : a.Ui[j][i];
        sr += ur * src[2 * j] + ui * src[2 * j + 1];
        si += ur * src[2 * j + 1] - ui * src[2 * j];
      }
      yr[i] = sr;
      yi[i] = si;
    }
  }
  double* outp = a.out + b * (long long)a.T_out * 2 * D;
  int out_slot = 0;
  auto emit = [&](const double* vr, const double* vi) {
    double* o = outp + (long long)out_slot * 2 * D;
    out_slot++;
#pragma unroll
    for (int i = 0; i < D; i++) {
      double sr = 0, si = 0;
#pragma unroll
      for (int j = 0; j < D; j++) {
        const double ur = a.identityU ? (i == j ? 1.0 : 0.0) : a.Ur[i][j];
        const double ui = a.identityU ? 0.0 : a.Ui[i][j];
        sr += ur * vr[j] - ui * vi[j];
        si += ur * vi[j] + ui * vr[j];
      }
      o[2 * i] = sr;
      o[2 * i + 1] = si;
    }
  };
  if (a.keep[0]) emit(yr, yi);

  double t = a.tpts[0];
  // k[0] = f(y0, t0)
  double kr[7][D], ki[7][D];
  dp_rhs<D, KA, TRI>(a, b, t, yr, yi, kr[0], ki[0]);

  // ---- initial step size (Hairer, Norsett, Wanner II.4; order 4)
  double dt;
  {
    double d0 = 0, d1 = 0;
    double sc[D];
#pragma unroll
    for (int i = 0; i < D; i++) {
      sc[i] = a.atol + sqrt(yr[i] * yr[i] + yi[i] * yi[i]) * a.rtol;
      d0 += (yr[i] * yr[i] + yi[i] * yi[i]) / (sc[i] * sc[i]);
      d1 += (kr[0][i] * kr[0][i] + ki[0][i] * ki[0][i]) / (sc[i] * sc[i]);
    }
    d0 = sqrt(d0);
    d1 = sqrt(d1);
    const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    double tr[D], ti[D], fr[D], fi[D];
#pragma unroll
    for (int i = 0; i < D; i++) {
      tr[i] = yr[i] + h0 * kr[0][i];
      ti[i] = yi[i] + h0 * ki[0][i];
    }
    dp_rhs<D, KA, TRI>(a, b, t + h0, tr, ti, fr, fi);
    double d2 = 0;
#pragma unroll
    for (int i = 0; i < D; i++) {
      const double er = fr[i] - kr[0][i], ei = fi[i] - ki[0][i];
      d2 += (er * er + ei * ei) / (sc[i] * sc[i]);
    }
    d2 = sqrt(d2) / h0;
    const double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / (d1 + d2), 1.0 / 5.0);
    dt = fmin(100.0 * h0, h1);
    dt = fmin(fmax(dt, 0.0), a.hmax);
  }

  long long n_att = 0, n_acc = 0;
  int status = 0;
  int next = 1;  // next merged time point to produce
  long long i_target = 0;
  double target = a.tpts[1 < a.T ? 1 : 0];  // a.tpts[next], reloaded only when `next` moves (not a load per attempt)
  while (next < a.T) {
    if (!(t < target)) {
      // cannot happen before the first step (t0 < t1); guard against equal points
      next++;
      if (next < a.T) target = a.tpts[next];
      continue;
    }
    if (i_target >= a.mxstep) {
      status = 1;
      break;
    }
    if (!(dt > 0.0)) {
      status = 2;
      break;
    }
    // ---- one attempted step.  The t-only part of the six new stages (five distinct times: stages 5 and 6 are both
    // at t + dt) does not depend on the state: the five envelope look-ups are issued first (their latency hides behind
    // the sin/cos work), then all 5 x NPH sin/cos pairs as one block of independent chains; the per-stage coefficient
    // sets are assembled between the stages, off the critical path.  (A rolled loop over the five times kept the code
    // small but serialised them: 3.9 ms against 2.95 ms at cfg 5.)
    {
      constexpr int NU = 5 * NPH;
      double2 envs[5][KA];
#pragma unroll
      for (int s = 0; s < 5; s++) dp_env<KA>(a, b, fma(dt, DP_ALPHA[s], t), envs[s]);
      double usn[NU], ucs[NU];
      {
        double ux[NU];
#pragma unroll
        for (int u = 0; u < NU; u++) ux[u] = a.phase_freq[u % NPH] * fma(dt, DP_ALPHA[u / NPH], t);
        dp_sincos_n<NU>(dp_fast_range(a, t) && dp_fast_range(a, t + dt), ux, usn, ucs);
      }
      DpCoef<D> cf;
#pragma unroll
      for (int s = 1; s < 7; s++) {
        if (s < 6) {
          dp_build<D, KA, TRI>(a, envs[s - 1], usn + (s - 1) * NPH, ucs + (s - 1) * NPH, cf);
        }
        double tr[D], ti[D];
#pragma unroll
        for (int i = 0; i < D; i++) {
          double ar = 0, ai = 0;
#pragma unroll
          for (int q = 0; q < s; q++) {
            ar += DP_BETA[s - 1][q] * kr[q][i];
            ai += DP_BETA[s - 1][q] * ki[q][i];
          }
          tr[i] = yr[i] + dt * ar;
          ti[i] = yi[i] + dt * ai;
        }
        dp_apply<D, TRI>(a.emask, cf, tr, ti, kr[s], ki[s]);
      }
    }
    double y1r[D], y1i[D];
    double ratio = 0;
#pragma unroll
    for (int i = 0; i < D; i++) {
      double sr = 0, si = 0, er = 0, ei = 0;
#pragma unroll
      for (int q = 0; q < 7; q++) {
        const double cs = (q < 6) ? DP_BETA[5][q] : 0.0;
        sr += cs * kr[q][i];
        si += cs * ki[q][i];
        er += DP_CERR[q] * kr[q][i];
        ei += DP_CERR[q] * ki[q][i];
      }
      y1r[i] = dt * sr + yr[i];
      y1i[i] = dt * si + yi[i];
      er *= dt;
      ei *= dt;
      ratio += dp_err_term(er * er + ei * ei, fmax(yr[i] * yr[i] + yi[i] * yi[i], y1r[i] * y1r[i] + y1i[i] * y1i[i]), a.atol, a.rtol);
    }
    ratio *= 1.0 / D;  // MEAN SQUARE error ratio from here on: accept iff <= 1 (dp_step_factor)
    double new_dt = dt * dp_step_factor(ratio);
    new_dt = fmin(fmax(new_dt, 0.0), a.hmax);
    n_att++;
    i_target++;
    if (ratio <= 1.0) {
      n_acc++;
      const double t_new = t + dt;
      // outputs inside (t, t_new]: 4th-order interpolant fitted to this step
      if (t_new >= target) {
        double pa_r[D], pa_i[D], pb_r[D], pb_i[D], pc_r[D], pc_i[D], pd_r[D], pd_i[D];
#pragma unroll
        for (int i = 0; i < D; i++) {
          double mr = 0, mi = 0;
#pragma unroll
          for (int q = 0; q < 7; q++) {
            mr += DP_CMID[q] * kr[q][i];
            mi += DP_CMID[q] * ki[q][i];
          }
          const double ymr = yr[i] + dt * mr, ymi = yi[i] + dt * mi;
          const double d0r = dt * kr[0][i], d0i = dt * ki[0][i], d1r = dt * kr[6][i], d1i = dt * ki[6][i];
          pa_r[i] = -2.0 * d0r + 2.0 * d1r - 8.0 * yr[i] - 8.0 * y1r[i] + 16.0 * ymr;
          pa_i[i] = -2.0 * d0i + 2.0 * d1i - 8.0 * yi[i] - 8.0 * y1i[i] + 16.0 * ymi;
          pb_r[i] = 5.0 * d0r - 3.0 * d1r + 18.0 * yr[i] + 14.0 * y1r[i] - 32.0 * ymr;
          pb_i[i] = 5.0 * d0i - 3.0 * d1i + 18.0 * yi[i] + 14.0 * y1i[i] - 32.0 * ymi;
          pc_r[i] = -4.0 * d0r + d1r - 11.0 * yr[i] - 5.0 * y1r[i] + 16.0 * ymr;
          pc_i[i] = -4.0 * d0i + d1i - 11.0 * yi[i] - 5.0 * y1i[i] + 16.0 * ymi;
          pd_r[i] = d0r;
          pd_i[i] = d0i;
        }
        while (next < a.T && a.tpts[next] <= t_new) {
          const double rel = (a.tpts[next] - t) / (t_new - t);
          if (a.keep[next]) {
            double vr[D], vi[D];
#pragma unroll
            for (int i = 0; i < D; i++) {
              vr[i] = (((pa_r[i] * rel + pb_r[i]) * rel + pc_r[i]) * rel + pd_r[i]) * rel + yr[i];
              vi[i] = (((pa_i[i] * rel + pb_i[i]) * rel + pc_i[i]) * rel + pd_i[i]) * rel + yi[i];
            }
            emit(vr, vi);
          }
          next++;
          i_target = 0;
        }
        if (next < a.T) target = a.tpts[next];
      }
      t = t_new;
#pragma unroll
      for (int i = 0; i < D; i++) {
        yr[i] = y1r[i];
        yi[i] = y1i[i];
        kr[0][i] = kr[6][i];  // FSAL
        ki[0][i] = ki[6][i];
      }
    }
    dt = new_dt;
  }
  // failed trajectories: fill the remaining outputs with NaN so nobody mistakes them for results
  if (status != 0) {
    double nanv[D];
#pragma unroll
    for (int i = 0; i < D; i++) nanv[i] = nan("");
    for (; next < a.T; next++)
      if (a.keep[next]) emit(nanv, nanv);
  }
  if (a.nsteps) {
    a.nsteps[2 * b] = n_att;
    a.nsteps[2 * b + 1] = n_acc;
  }
  if (a.status) a.status[b] = status;
}

int casq_prepare_slots(const casq_model* m, int n_slots, const casq_pulse_slot* slots, DevSlots* ds, int* active,
                       int* n_active, int* n_total);
bool casq_dopri5_general_supported(const casq_model* m, int n_active);
int casq_dopri5_general(casq_model* m, int KA, const int* active, const DevSlots& slots, int n_total, long long B,
                        const double* params_dev, const double* y0_dev, int y0_batched, double rtol, double atol, double hmax,
                        long long mxstep, const std::vector<double>& pts, const std::vector<char>& keep, int T_out,
                        double* out_dev, long long* nsteps_dev, int* status_dev, cudaStream_t st);

extern "C" int casq_solve_sv_dopri5(casq_model* m, int n_slots, const casq_pulse_slot* slots, int64_t B,
                                    const double* params_dev, const double* y0_dev, int y0_batched, double t0, double t1,
                                    double rtol, double atol, double hmax, int64_t mxstep, int T, const double* t_eval,
                                    double* out_dev, int64_t* nsteps_dev, int32_t* status_dev, void* stream) {
  if (!m) CASQ_FAIL(CASQ_ERR_INVALID, "casq_solve_sv_dopri5: NULL model");
  if (B < 0) CASQ_FAIL(CASQ_ERR_INVALID, "casq_solve_sv_dopri5: negative batch size");
  if (B > 0 && (!y0_dev || !out_dev || (n_slots > 0 && !params_dev))) CASQ_FAIL(CASQ_ERR_INVALID, "casq_solve_sv_dopri5: NULL buffer");
  if (!(t1 > t0)) CASQ_FAIL(CASQ_ERR_INVALID, "casq_solve_sv_dopri5: need t1 > t0");
  if (!(rtol > 0) || !(atol > 0)) CASQ_FAIL(CASQ_ERR_INVALID, "casq_solve_sv_dopri5: rtol and atol must be > 0");
  DevSlots ds;
  int active[CASQ_MAX_SLOTS], KA = 0, n_total = 0;
  int rc = casq_prepare_slots(m, n_slots, slots, &ds, active, &KA, &n_total);
  if (rc != CASQ_OK) return rc;
  if (B == 0) return CASQ_OK;
  if (KA == 0) {
    if (m->K < 1) CASQ_FAIL(CASQ_ERR_UNSUPPORTED, "model without channels");
    active[0] = 0;
    KA = 1;
  }
  const int d = m->d;
  const bool small_path = d >= 2 && d <= DP_MAXD && KA <= DP_MAXK;
  if (!small_path && !casq_dopri5_general_supported(m, KA))
    CASQ_FAIL(CASQ_ERR_UNSUPPORTED, "casq_solve_sv_dopri5: dim %d with %d active channels is outside the built kernels (dim <= 256, <= 4 channels)", d, KA);
  cudaStream_t st = (cudaStream_t)stream;
  CASQ_CUDA(cudaSetDevice(m->device));

  // merged output grid (same trimming rule as the fixed-step path)
  std::vector<double> pts;
  std::vector<char> keep;
  if (!t_eval) {
    if (T != 2) CASQ_FAIL(CASQ_ERR_INVALID, "t_eval is NULL: T must be 2 (t0 and t1), got %d", T);
    pts = {t0, t1};
    keep = {1, 1};
  } else {
    if (T < 1) CASQ_FAIL(CASQ_ERR_INVALID, "t_eval needs at least one point");
    for (int i = 1; i < T; i++)
      if (!(t_eval[i] > t_eval[i - 1])) CASQ_FAIL(CASQ_ERR_INVALID, "t_eval must be strictly increasing");
    if (t_eval[0] < t0 || t_eval[T - 1] > t1) CASQ_FAIL(CASQ_ERR_INVALID, "t_eval must lie inside [t0, t1]");
    if (t_eval[0] != t0) {
      pts.push_back(t0);
      keep.push_back(0);
    }
    for (int i = 0; i < T; i++) {
      pts.push_back(t_eval[i]);
      keep.push_back(1);
    }
    if (t_eval[T - 1] != t1) {
      pts.push_back(t1);
      keep.push_back(0);
    }
  }
  const int TM = (int)pts.size();
  if (!small_path)
    return casq_dopri5_general(m, KA, active, ds, n_total, B, params_dev, y0_dev, y0_batched, rtol, atol,
                               (hmax > 0) ? hmax : INFINITY, (mxstep > 0) ? (long long)mxstep : (1LL << 62), pts, keep, T, out_dev,
                               (long long*)nsteps_dev, status_dev, st);
  if ((rc = m->tab_times.ensure(TM * sizeof(double)))) return rc;
  if ((rc = m->consts.ensure(TM))) return rc;
  m->tab_key.clear();  // the fixed-step table cache shares these buffers
  CASQ_CUDA(cudaMemcpyAsync(m->tab_times.p, pts.data(), TM * sizeof(double), cudaMemcpyHostToDevice, st));
  CASQ_CUDA(cudaMemcpyAsync(m->consts.p, keep.data(), TM, cudaMemcpyHostToDevice, st));

  Dopri5Args a;
  memset(&a, 0, sizeof(a));
  a.B = B;
  a.T = TM;
  a.tpts = (const double*)m->tab_times.p;
  a.keep = (const char*)m->consts.p;
  a.T_out = T;
  a.params = params_dev;
  a.y0 = y0_dev;
  a.y0_batched = y0_batched;
  a.out = out_dev;
  a.nsteps = (long long*)nsteps_dev;
  a.status = status_dev;
  a.n_total = n_total;
  a.dt_sample = m->dt;
  a.inv_dt_sample = 1.0 / m->dt;
  a.rtol = rtol;
  a.atol = atol;
  a.hmax = (hmax > 0) ? hmax : INFINITY;
  a.mxstep = (mxstep > 0) ? mxstep : (1LL << 62);
  a.rwa = m->rwa;
  for (int k = 0; k < KA; k++) a.two_pi_nu[k] = 2.0 * M_PI * m->nu[active[k]];
  int e = 0;
  for (int i = 0; i < d; i++) {
    a.Sd[i] = m->S[i * d + i].real();
    if (a.Sd[i] != 0.0) a.has_diag = 1;
    for (int k = 0; k < KA; k++) {
      cplx p = m->P[((size_t)active[k] * d + i) * d + i];
      a.Pdr[k][i] = p.real();
      a.Pdi[k][i] = p.imag();
      if (std::abs(p) != 0.0) a.has_diag = 1;
    }
    for (int j = 0; j < d; j++) {
      a.Ur[i][j] = m->U[i * d + j].real();
      a.Ui[i][j] = m->U[i * d + j].imag();
    }
    for (int j = i + 1; j < d; j++) {
      a.bohr[e] = m->fe[i] - m->fe[j];
      cplx s = m->S[i * d + j];
      a.Sr[e] = s.real();
      a.Si[e] = s.imag();
      bool any = std::abs(s) != 0.0;
      for (int k = 0; k < KA; k++) {
        cplx p = m->P[((size_t)active[k] * d + i) * d + j], n = m->N[((size_t)active[k] * d + i) * d + j];
        a.Pr[k][e] = p.real();
        a.Pi[k][e] = p.imag();
        a.Nr[k][e] = n.real();
        a.Ni[k][e] = n.imag();
        any = any || std::abs(p) != 0.0 || std::abs(n) != 0.0;
      }
      if (any) a.emask |= (1u << e);
      e++;
    }
  }
  a.identityU = m->U_is_identity ? 1 : 0;
  // ladder-only models take the TRI variants (D = 2 has nothing else)
  unsigned ladder = 0;
  for (int i = 0; i + 1 < d; i++) ladder |= 1u << (i * (2 * d - i - 1) / 2);
  const bool tri = (a.emask & ~ladder) == 0;
  {
    int n = 0;
    for (int k = 0; k < KA; k++) a.phase_freq[n++] = a.two_pi_nu[k];
    for (int k = 0; k < KA; k++) a.freq_max = std::max(a.freq_max, std::fabs(a.two_pi_nu[k]));
    for (int e2 = 0; e2 < d * (d - 1) / 2; e2++) a.freq_max = std::max(a.freq_max, std::fabs(a.bohr[e2]));
    const int E = d * (d - 1) / 2;
    for (int e2 = 0; e2 < E; e2++) {
      if (tri && !(ladder & (1u << e2))) continue;
      a.phase_freq[n++] = (a.emask & (1u << e2)) ? a.bohr[e2] : 0.0;
    }
  }
  // the batch is walked in windows that bound the envelope table (one window up to 2.6e5 points at cfg 5's 257 samples)
  const long long Bw = casq_env_table_window(KA, n_total, B);
  if ((rc = m->work2.ensure((size_t)KA * std::max(n_total, 1) * (size_t)Bw * sizeof(double2)))) return rc;
  a.env_tab = (const double2*)m->work2.p;
  const int block = 64;
  for (long long b0 = 0; b0 < B; b0 += Bw) {
    const long long Bc = std::min(Bw, (long long)B - b0);
    CASQ_CUDA(casq_launch_env_table<DP_MAXK>(KA, ds, params_dev, B, b0, Bc, n_total, (double2*)m->work2.p, st));
    a.B = Bc;
    a.y0 = y0_dev + (y0_batched ? b0 * 2 * d : 0);
    a.out = out_dev + b0 * (long long)T * 2 * d;
    a.nsteps = nsteps_dev ? (long long*)nsteps_dev + 2 * b0 : nullptr;
    a.status = status_dev ? status_dev + b0 : nullptr;
    bool launched = false;
    const unsigned grid = (unsigned)((Bc + block - 1) / block);
#define LAUNCH(D_, K_, T_)                                      \
  if (d == D_ && KA == K_ && tri == T_) {                      \
    sv_dopri5_kernel<D_, K_, T_><<<grid, block, 0, st>>>(a);   \
    launched = true;                                           \
  }
    LAUNCH(2, 1, true) LAUNCH(2, 2, true) LAUNCH(3, 1, true) LAUNCH(3, 2, true) LAUNCH(4, 1, true) LAUNCH(4, 2, true)
    LAUNCH(3, 1, false) LAUNCH(3, 2, false) LAUNCH(4, 1, false) LAUNCH(4, 2, false)
#undef LAUNCH
    if (!launched) CASQ_FAIL(CASQ_ERR_UNSUPPORTED, "casq_solve_sv_dopri5: no kernel variant for dim %d, %d channels", d, KA);
  }
  CASQ_CUDA(cudaGetLastError());
  return CASQ_OK;
}
