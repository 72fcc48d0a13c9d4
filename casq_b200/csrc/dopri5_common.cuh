// casq_b200: pieces shared by the two adaptive Dopri5 steppers (thread per trajectory, d <= 4; warp per trajectory,
// d <= 64): the sample index of a per-trajectory time and a branch-free fp64 sincos.
#pragma once
#include "common.cuh"

// numpy float floor-division a // b (b > 0), then clip to [-1, n_total]
static __device__ __noinline__ int dev_sample_index(double t, double dt, int n_total) {
  double mod = fmod(t, dt);
  double div = (t - mod) / dt;
  if (mod != 0.0 && (mod < 0)) div -= 1.0;
  double fl = 0.0;
  if (div != 0.0) {
    fl = floor(div);
    if (div - fl > 0.5) fl += 1.0;
  }
  if (fl < -1.0) fl = -1.0;
  if (fl > (double)n_total) fl = (double)n_total;
  return (int)fl;
}

// The same index without the iterative fmod and without the conversion unit: q = NEAREST integer to t / dt through the
// 2^52 + 2^51 constant (the integer is the low word of the sum), then the EXACT sign of the remainder t - q dt (one FMA:
// the rounded result has the sign of the exact one) says whether floor(t / dt) is q or q - 1.  With the nearest integer
// the remainder lies in [-dt/2, dt/2], so the ambiguous "remainder rounds to dt" case of a floor-based guess cannot
// occur.  Quotients beyond the int range take the reference path above.
__device__ __forceinline__ int dev_sample_index_fast(double t, double dt, double inv_dt, int n_total) {
  const double x = t * inv_dt;
  if (!(fabs(x) < 2.0e9)) return dev_sample_index(t, dt, n_total);
  const double magic = 6755399441055744.0;
  const double m = x + magic;
  const double q = m - magic;
  const double r = fma(-q, dt, t);
  const int qi = __double2loint(m) - (r < 0.0 ? 1 : 0);
  return min(max(qi, -1), n_total);
}

// Branch-free fp64 sincos for |x| < 1e5 (phases reach ~2e3 rad here): two-FMA Cody-Waite reduction by pi/2 and the
// fdlibm kernel polynomials on [-pi/4, pi/4] (< 1 ulp).  The library sincos hides a slow-path branch in every call,
// which keeps the compiler from overlapping the carrier and Bohr-phase evaluations of a stage; callers test the
// argument range once for all phases of a stage (DP_TRIG_MAX).  Coefficients live in constant memory so that every
// DFMA takes its constant straight from the bank (as 64-bit immediates each one cost two UMOV per use).
#define DP_TRIG_MAX 1.0e5
static __constant__ double DP_TRIG[16] = {6.36619772367581382433e-01, 1.5707963267948966, 6.123233995736766e-17, 0.0,
                                   1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,
                                   -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01,
                                   -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,
                                   2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02};
__device__ __forceinline__ void dp_sincos_fast(double x, double* sn, double* cs) {
  const double n = rint(x * DP_TRIG[0]);
  double r = fma(-n, DP_TRIG[1], x);
  r = fma(-n, DP_TRIG[2], r);
  const double z = r * r;
  double ps = fma(z, DP_TRIG[4], DP_TRIG[5]);
  ps = fma(z, ps, DP_TRIG[6]);
  ps = fma(z, ps, DP_TRIG[7]);
  ps = fma(z, ps, DP_TRIG[8]);
  ps = fma(z, ps, DP_TRIG[9]);
  const double sr = fma(z * r, ps, r);
  double pc = fma(z, DP_TRIG[10], DP_TRIG[11]);
  pc = fma(z, pc, DP_TRIG[12]);
  pc = fma(z, pc, DP_TRIG[13]);
  pc = fma(z, pc, DP_TRIG[14]);
  pc = fma(z, pc, DP_TRIG[15]);
  const double cr = fma(z * z, pc, fma(-0.5, z, 1.0));
  const int q = (int)n;
  const double s0 = (q & 1) ? cr : sr, c0 = (q & 1) ? sr : cr;
  // quadrant signs flipped in the sign bit directly (one LOP3 each; the ?: form cost a negate and two selects)
  *sn = __hiloint2double(__double2hiint(s0) ^ ((q & 2) << 30), __double2loint(s0));
  *cs = __hiloint2double(__double2hiint(c0) ^ (((q + 1) & 2) << 30), __double2loint(c0));
}


// Step-size factor of the controller, min(10, max(0.9 ratio^(-1/5), ratio < 1 ? 1 : 0.2)), from the MEAN SQUARE error
// ratio m = ratio^2 (so the caller needs no square root: ratio <= 1 iff m <= 1).  m^(-1/10) comes from an fp32 seed and
// two Newton steps on x^-10 = m (x <- x (1.1 - 0.1 m x^10), error 5.5 e^2 per step: 1e-5 -> 6e-10 -> 2e-18), ~25
// instructions where exp2(-0.2 log2(ratio)) took 155 and sat on the critical path of every step.  Outside [1e-14, 1e8]
// the result is clipped to 10 or 0.2 anyway, so m is clamped first (which also covers m = 0 and keeps the seed finite).
__device__ __forceinline__ double dp_step_factor(double m) {
  const double dfac = m < 1.0 ? 1.0 : 0.2;
  const double r = (m == m) ? fmin(fmax(m, 1e-14), 1e8) : 1e8;  // NaN error estimate: shrink like a huge one
  double x = (double)exp2f(-0.1f * __log2f((float)r));
#pragma unroll
  for (int it = 0; it < 2; it++) {
    const double x2 = x * x, x4 = x2 * x2, x8 = x4 * x4;
    x = x * fma(-0.1 * r, x8 * x2, 1.1);
  }
  return fmin(10.0, fmax(0.9 * x, dfac));
}

// (er^2 + ei^2) / tol^2 with tol = atol + rtol sqrt(a), a = max(|y0|^2, |y1|^2): the error-ratio term of one component.
// The IEEE sqrt and division (with their special-case branches) were ~50 instructions per component on the critical
// path of every step; here 1/sqrt(a) and 1/tol come from the hardware seeds (2^-20) and two Newton steps each, accurate to
// 1-2 ulp -- an error estimate needs no more.  tol >= atol > 0 always; a = 0 (or subnormal) gives sqrt = 0.
__device__ __forceinline__ double dp_err_term(double e2, double a, double atol, double rtol) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double ha = 0.5 * a;
  y = y * fma(-ha * y, y, 1.5);
  y = y * fma(-ha * y, y, 1.5);
  const double s = a > 1e-290 ? a * y : 0.0;
  const double tol = fma(rtol, s, atol);
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(tol));
  x = fma(x, fma(-tol, x, 1.0), x);
  x = fma(x, fma(-tol, x, 1.0), x);
  return e2 * (x * x);
}
