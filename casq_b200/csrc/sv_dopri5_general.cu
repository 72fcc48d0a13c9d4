// casq_b200: adaptive Dormand-Prince 5(4) for GENERAL Hilbert-space dimension (5..64): one WARP per trajectory
// (lane = row, two rows per lane above 32), per-warp step control, sparse-row matvec through shared memory.
//
// Completes ODESolverMethod.QISKIT_DYNAMICS_JAX_ODEINT (src/casq/backends/pulse_backend.py:65) and the
// scipy-named adaptive methods for coupled-transmon models behind QiskitPulseBackend.solve
// (src/casq/backends/qiskit/qiskit_pulse_backend.py:158-162): same controller as sv_dopri5.cu (the published
// jax.experimental.ode.odeint algorithm), same right-hand side as sv_general.cu, except that stage times differ
// per trajectory so p_i = e^{i e_i t} and the carriers are fp64 sincos per stage.  A warp per trajectory keeps the
// accept/reject decision warp-uniform (no intra-warp divergence), at the price of idle lanes for d < 32.
#include <algorithm>
#include <cstring>

#include "common.cuh"
#include "envelope.cuh"
#include "dopri5_common.cuh"

#define DG_MAXK 4
#define DG_WARPS 4

struct DopriGenArgs {
  long long B;
  int T, T_out, d, KA, width, rwa;
  const double* tpts;
  const char* keep;
  const double* params;
  const double* y0;
  int y0_batched;
  double* out;
  long long* nsteps;
  int* status;
  const double2* env_tab;  // [KA][n_total][B] envelope samples (casq_env_table_kernel)
  int n_total;
  double dt_sample, inv_dt_sample, rtol, atol, hmax;
  double max_abs_freq;  // largest |carrier or frame frequency| (rad/s): decides the branch-free sincos per call
  long long mxstep;
  double two_pi_nu[DG_MAXK];
  const double* fe;       // [d] frame eigenvalues
  const int* ell_col;     // [width][d]
  const double2* ell_S;   // [width][d]
  const double2* ell_P;   // [KA][width][d]
  const double2* ell_N;   // [KA][width][d]
  int identityU;
  const double2* U;       // [d][d]
};

__device__ __forceinline__ double dg_warp_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__constant__ double DG_ALPHA[6] = {1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1.0, 1.0};
__constant__ double DG_BETA[6][6] = {
    {1.0 / 5, 0, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656, 0},
    {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84}};
__constant__ double DG_CERR[7] = {35.0 / 384 - 1951.0 / 21600, 0, 500.0 / 1113 - 22642.0 / 50085, 125.0 / 192 - 451.0 / 720,
                                  -2187.0 / 6784 - -12231.0 / 42400, 11.0 / 84 - 649.0 / 6300, -1.0 / 60};
__constant__ double DG_CMID[7] = {6025192743.0 / 30085553152.0 / 2, 0, 51252292925.0 / 65400821598.0 / 2,
                                  -2691868925.0 / 45128329728.0 / 2, 187940372067.0 / 1594534317056.0 / 2,
                                  -1776094331.0 / 19743644256.0 / 2, 11237099.0 / 235043384.0 / 2};

template <int R, bool RWA>
__global__ void __launch_bounds__(32 * DG_WARPS) sv_dopri5_general_kernel(const DopriGenArgs a) {
  const int d = a.d, KA = a.KA, width = a.width;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int* s_col = reinterpret_cast<int*>(smem_raw);
  size_t off = ((size_t)width * d * sizeof(int) + 15) & ~(size_t)15;
  double2* s_S = reinterpret_cast<double2*>(smem_raw + off);
  double2* s_P = s_S + (size_t)width * d;
  double2* s_N = s_P + (size_t)KA * width * d;
  double2* s_z = (RWA ? s_N + (size_t)KA * width * d : s_N);
  for (int i = threadIdx.x; i < width * d; i += blockDim.x) {
    s_col[i] = a.ell_col[i];
    s_S[i] = a.ell_S[i];
  }
  for (int i = threadIdx.x; i < KA * width * d; i += blockDim.x) {
    s_P[i] = a.ell_P[i];
    if (RWA) s_N[i] = a.ell_N[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long b = (long long)blockIdx.x * DG_WARPS + warp;
  if (b >= a.B) return;
  double2* zb = s_z + (size_t)warp * (R * 32);
  bool own[R];
  double fe[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    own[r] = (lane + 32 * r) < d;
    fe[r] = own[r] ? a.fe[lane + 32 * r] : 0.0;
  }

  // k = -i p (A (conj(p) v)) at time t
  auto rhs = [&](double t, const double* vr, const double* vi, double* kr, double* ki) {
    // envelope samples from the per-solve table (casq_env_table_kernel); phases below |x| = 1e5 with the branch-free
    // sincos (one warp-uniform range test per call: t is the same for all lanes)
    const int gidx = dev_sample_index_fast(t, a.dt_sample, a.inv_dt_sample, a.n_total);
    const bool inside = gidx >= 0 && gidx < a.n_total;
    const bool fast_trig = a.max_abs_freq * fabs(t) < DP_TRIG_MAX;
    double zr[DG_MAXK], zi[DG_MAXK];
#pragma unroll
    for (int k = 0; k < DG_MAXK; k++) {
      zr[k] = zi[k] = 0.0;
      if (k < KA && inside) {
        const double2 env = __ldg(a.env_tab + ((long long)k * a.n_total + gidx) * a.B + b);
        double s, c;
        if (fast_trig)
          dp_sincos_fast(a.two_pi_nu[k] * t, &s, &c);
        else
          sincos(a.two_pi_nu[k] * t, &s, &c);
        zr[k] = env.x * c - env.y * s;
        zi[k] = env.x * s + env.y * c;
      }
    }
    double pr[R], pi[R];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < R; r++) {
      pr[r] = 1.0;
      pi[r] = 0.0;
      if (own[r]) {
        if (fast_trig)
          dp_sincos_fast(fe[r] * t, &pi[r], &pr[r]);
        else
          sincos(fe[r] * t, &pi[r], &pr[r]);
        zb[lane + 32 * r] = make_double2(pr[r] * vr[r] + pi[r] * vi[r], pr[r] * vi[r] - pi[r] * vr[r]);
      }
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int i = lane + 32 * r;
      double wr = 0.0, wi = 0.0;
      if (own[r]) {
        for (int q = 0; q < width; q++) {
          const int e = q * d + i;
          double2 A = s_S[e];
#pragma unroll
          for (int k = 0; k < DG_MAXK; k++) {
            if (k < KA) {
              const double2 P = s_P[(size_t)k * width * d + e];
              if (RWA) {
                const double2 N = s_N[(size_t)k * width * d + e];
                A.x += zr[k] * (P.x + N.x) - zi[k] * (P.y - N.y);
                A.y += zr[k] * (P.y + N.y) + zi[k] * (P.x - N.x);
              } else {
                A.x = fma(zr[k], P.x, A.x);
                A.y = fma(zr[k], P.y, A.y);
              }
            }
          }
          const double2 z = zb[s_col[e]];
          wr += A.x * z.x - A.y * z.y;
          wi += A.x * z.y + A.y * z.x;
        }
      }
      const double tr = pr[r] * wr - pi[r] * wi, ti = pr[r] * wi + pi[r] * wr;
      kr[r] = ti;
      ki[r] = -tr;
    }
  };

  // y <- U^+ y0
  double yr[R], yi[R];
  {
    const double* src = a.y0 + (a.y0_batched ? b * 2 * d : 0);
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int i = lane + 32 * r;
      double sr = 0, si = 0;
      if (own[r]) {
        if (a.identityU) {
          sr = src[2 * i];
          si = src[2 * i + 1];
        } else {
          for (int j = 0; j < d; j++) {
            const double2 u = a.U[j * d + i];
            sr += u.x * src[2 * j] + u.y * src[2 * j + 1];
            si += u.x * src[2 * j + 1] - u.y * src[2 * j];
          }
        }
      }
      yr[r] = sr;
      yi[r] = si;
    }
  }
  double* outp = a.out + b * (long long)a.T_out * 2 * d;
  int out_slot = 0;
  auto emit = [&](const double* vr, const double* vi, bool nan_fill) {
    double* o = outp + (long long)out_slot * 2 * d;
    if (a.identityU || nan_fill) {
#pragma unroll
      for (int r = 0; r < R; r++)
        if (own[r]) {
          o[2 * (lane + 32 * r)] = nan_fill ? nan("") : vr[r];
          o[2 * (lane + 32 * r) + 1] = nan_fill ? nan("") : vi[r];
        }
    } else {
      __syncwarp();
#pragma unroll
      for (int r = 0; r < R; r++)
        if (own[r]) zb[lane + 32 * r] = make_double2(vr[r], vi[r]);
      __syncwarp();
#pragma unroll
      for (int r = 0; r < R; r++) {
        const int i = lane + 32 * r;
        if (own[r]) {
          double sr = 0, si = 0;
          for (int j = 0; j < d; j++) {
            const double2 u = a.U[i * d + j], v = zb[j];
            sr += u.x * v.x - u.y * v.y;
            si += u.x * v.y + u.y * v.x;
          }
          o[2 * i] = sr;
          o[2 * i + 1] = si;
        }
      }
      __syncwarp();
    }
    out_slot++;
  };
  if (a.keep[0]) emit(yr, yi, false);

  double t = a.tpts[0];
  double kr[7][R], ki[7][R];
  rhs(t, yr, yi, kr[0], ki[0]);

  // ---- initial step size (Hairer, Norsett, Wanner II.4; order 4)
  double dt;
  {
    double d0 = 0, d1 = 0, sc[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      sc[r] = a.atol + hypot(yr[r], yi[r]) * a.rtol;
      if (own[r]) {
        d0 += (yr[r] * yr[r] + yi[r] * yi[r]) / (sc[r] * sc[r]);
        d1 += (kr[0][r] * kr[0][r] + ki[0][r] * ki[0][r]) / (sc[r] * sc[r]);
      }
    }
    d0 = sqrt(dg_warp_sum(d0));
    d1 = sqrt(dg_warp_sum(d1));
    const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    double tr[R], ti[R], fr[R], fi[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      tr[r] = yr[r] + h0 * kr[0][r];
      ti[r] = yi[r] + h0 * ki[0][r];
    }
    rhs(t + h0, tr, ti, fr, fi);
    double d2 = 0;
#pragma unroll
    for (int r = 0; r < R; r++)
      if (own[r]) {
        const double er = fr[r] - kr[0][r], ei = fi[r] - ki[0][r];
        d2 += (er * er + ei * ei) / (sc[r] * sc[r]);
      }
    d2 = sqrt(dg_warp_sum(d2)) / h0;
    const double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / (d1 + d2), 1.0 / 5.0);
    dt = fmin(fmax(fmin(100.0 * h0, h1), 0.0), a.hmax);
  }

  long long n_att = 0, n_acc = 0, i_target = 0;
  int status = 0, next = 1;
  while (next < a.T) {
    const double target = a.tpts[next];
    if (!(t < target)) {
      next++;
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
#pragma unroll
    for (int s = 1; s < 7; s++) {
      double tr[R], ti[R];
#pragma unroll
      for (int r = 0; r < R; r++) {
        double ar = 0, ai = 0;
#pragma unroll
        for (int q = 0; q < s; q++) {
          ar += DG_BETA[s - 1][q] * kr[q][r];
          ai += DG_BETA[s - 1][q] * ki[q][r];
        }
        tr[r] = yr[r] + dt * ar;
        ti[r] = yi[r] + dt * ai;
      }
      rhs(t + dt * DG_ALPHA[s - 1], tr, ti, kr[s], ki[s]);
    }
    double y1r[R], y1i[R], ratio = 0;
#pragma unroll
    for (int r = 0; r < R; r++) {
      double sr = 0, si = 0, er = 0, ei = 0;
#pragma unroll
      for (int q = 0; q < 7; q++) {
        const double cs = (q < 6) ? DG_BETA[5][q] : 0.0;
        sr += cs * kr[q][r];
        si += cs * ki[q][r];
        er += DG_CERR[q] * kr[q][r];
        ei += DG_CERR[q] * ki[q][r];
      }
      y1r[r] = dt * sr + yr[r];
      y1i[r] = dt * si + yi[r];
      er *= dt;
      ei *= dt;
      if (own[r])
        ratio += dp_err_term(er * er + ei * ei, fmax(yr[r] * yr[r] + yi[r] * yi[r], y1r[r] * y1r[r] + y1i[r] * y1i[r]), a.atol, a.rtol);
    }
    ratio = dg_warp_sum(ratio) / d;  // MEAN SQUARE error ratio from here on: accept iff <= 1 (dp_step_factor)
    double new_dt = dt * dp_step_factor(ratio);
    new_dt = fmin(fmax(new_dt, 0.0), a.hmax);
    n_att++;
    i_target++;
    if (ratio <= 1.0) {
      n_acc++;
      const double t_new = t + dt;
      if (t_new >= target) {
        double pa_r[R], pa_i[R], pb_r[R], pb_i[R], pc_r[R], pc_i[R], pd_r[R], pd_i[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
          double mr = 0, mi = 0;
#pragma unroll
          for (int q = 0; q < 7; q++) {
            mr += DG_CMID[q] * kr[q][r];
            mi += DG_CMID[q] * ki[q][r];
          }
          const double ymr = yr[r] + dt * mr, ymi = yi[r] + dt * mi;
          const double d0r = dt * kr[0][r], d0i = dt * ki[0][r], d1r = dt * kr[6][r], d1i = dt * ki[6][r];
          pa_r[r] = -2.0 * d0r + 2.0 * d1r - 8.0 * yr[r] - 8.0 * y1r[r] + 16.0 * ymr;
          pa_i[r] = -2.0 * d0i + 2.0 * d1i - 8.0 * yi[r] - 8.0 * y1i[r] + 16.0 * ymi;
          pb_r[r] = 5.0 * d0r - 3.0 * d1r + 18.0 * yr[r] + 14.0 * y1r[r] - 32.0 * ymr;
          pb_i[r] = 5.0 * d0i - 3.0 * d1i + 18.0 * yi[r] + 14.0 * y1i[r] - 32.0 * ymi;
          pc_r[r] = -4.0 * d0r + d1r - 11.0 * yr[r] - 5.0 * y1r[r] + 16.0 * ymr;
          pc_i[r] = -4.0 * d0i + d1i - 11.0 * yi[r] - 5.0 * y1i[r] + 16.0 * ymi;
          pd_r[r] = d0r;
          pd_i[r] = d0i;
        }
        while (next < a.T && a.tpts[next] <= t_new) {
          const double rel = (a.tpts[next] - t) / (t_new - t);
          if (a.keep[next]) {
            double vr[R], vi[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
              vr[r] = (((pa_r[r] * rel + pb_r[r]) * rel + pc_r[r]) * rel + pd_r[r]) * rel + yr[r];
              vi[r] = (((pa_i[r] * rel + pb_i[r]) * rel + pc_i[r]) * rel + pd_i[r]) * rel + yi[r];
            }
            emit(vr, vi, false);
          }
          next++;
          i_target = 0;
        }
      }
      t = t_new;
#pragma unroll
      for (int r = 0; r < R; r++) {
        yr[r] = y1r[r];
        yi[r] = y1i[r];
        kr[0][r] = kr[6][r];
        ki[0][r] = ki[6][r];
      }
    }
    dt = new_dt;
  }
  if (status != 0)
    for (; next < a.T; next++)
      if (a.keep[next]) emit(yr, yi, true);
  if (lane == 0) {
    if (a.nsteps) {
      a.nsteps[2 * b] = n_att;
      a.nsteps[2 * b + 1] = n_acc;
    }
    if (a.status) a.status[b] = status;
  }
}

bool casq_dopri5_general_supported(const casq_model* m, int n_active) {
  return m->d >= 2 && m->d <= 256 && n_active >= 1 && n_active <= DG_MAXK;
}

int casq_dopri5_general(casq_model* m, int KA, const int* active, const DevSlots& slots, int n_total, long long B,
                        const double* params_dev, const double* y0_dev, int y0_batched, double rtol, double atol, double hmax,
                        long long mxstep, const std::vector<double>& pts, const std::vector<char>& keep, int T_out,
                        double* out_dev, long long* nsteps_dev, int* status_dev, cudaStream_t st) {
  const int d = m->d;
  std::vector<std::vector<int>> cols(d);
  int width = 0;
  for (int i = 0; i < d; i++) {
    for (int j = 0; j < d; j++) {
      bool nz = std::abs(m->S[i * d + j]) != 0.0;
      for (int k = 0; k < KA && !nz; k++)
        nz = std::abs(m->P[((size_t)active[k] * d + i) * d + j]) != 0.0 || std::abs(m->N[((size_t)active[k] * d + i) * d + j]) != 0.0;
      if (nz) cols[i].push_back(j);
    }
    width = std::max(width, (int)cols[i].size());
  }
  if (width == 0) width = 1;
  const size_t ne = (size_t)width * d;
  std::vector<int> h_col(ne);
  std::vector<cplx> h_S(ne, 0.0), h_P((size_t)KA * ne, 0.0), h_N((size_t)KA * ne, 0.0);
  for (int i = 0; i < d; i++)
    for (int q = 0; q < width; q++) {
      const size_t e = (size_t)q * d + i;
      if (q < (int)cols[i].size()) {
        const int j = cols[i][q];
        h_col[e] = j;
        h_S[e] = m->S[i * d + j];
        for (int k = 0; k < KA; k++) {
          cplx p = m->P[((size_t)active[k] * d + i) * d + j], n = m->N[((size_t)active[k] * d + i) * d + j];
          h_P[(size_t)k * ne + e] = m->rwa ? p : 2.0 * p;
          h_N[(size_t)k * ne + e] = n;
        }
      } else {
        h_col[e] = i;
      }
    }
  // rows of the state per lane: 1, 2, 4 or 8.  From 4 up the seven stage vectors no longer fit the register file and
  // live in (L1-cached) local memory: the large models (4 and 5 transmons) run, at a lower rate per flop.
  const int R = d <= 32 ? 1 : (d <= 64 ? 2 : (d <= 128 ? 4 : 8));
  const size_t smem = (((size_t)ne * sizeof(int) + 15) & ~(size_t)15) + ne * 16 * (1 + (size_t)KA * (m->rwa ? 2 : 1)) +
                      (size_t)DG_WARPS * R * 32 * 16;
  if (smem > 200 * 1024)
    CASQ_FAIL(CASQ_ERR_UNSUPPORTED, "general Dopri5: operators too dense for shared memory (dim %d, row width %d, %d channels): large models need a diagonal rotating frame (rotating_frame = diag(static_operator), casq's from_backend default)", d, width, KA);
  const int TM = (int)pts.size();
  const size_t ell_bytes = ne * sizeof(int) + 16 + ne * 16 * (1 + 2 * (size_t)KA);
  int rc;
  m->tab_key.clear();
  if ((rc = m->tab_times.ensure(TM * sizeof(double)))) return rc;
  if ((rc = m->consts.ensure(TM))) return rc;
  if ((rc = m->work0.ensure((size_t)d * sizeof(double)))) return rc;
  if ((rc = m->work1.ensure(ell_bytes))) return rc;
  if ((rc = m->work2.ensure((size_t)d * d * 16))) return rc;
  CASQ_CUDA(cudaMemcpyAsync(m->tab_times.p, pts.data(), TM * sizeof(double), cudaMemcpyHostToDevice, st));
  CASQ_CUDA(cudaMemcpyAsync(m->consts.p, keep.data(), TM, cudaMemcpyHostToDevice, st));
  CASQ_CUDA(cudaMemcpyAsync(m->work0.p, m->fe.data(), (size_t)d * sizeof(double), cudaMemcpyHostToDevice, st));
  char* ell = (char*)m->work1.p;
  const size_t off_S = (ne * sizeof(int) + 15) & ~(size_t)15;
  CASQ_CUDA(cudaMemcpyAsync(ell, h_col.data(), ne * sizeof(int), cudaMemcpyHostToDevice, st));
  CASQ_CUDA(cudaMemcpyAsync(ell + off_S, h_S.data(), ne * 16, cudaMemcpyHostToDevice, st));
  CASQ_CUDA(cudaMemcpyAsync(ell + off_S + ne * 16, h_P.data(), (size_t)KA * ne * 16, cudaMemcpyHostToDevice, st));
  CASQ_CUDA(cudaMemcpyAsync(ell + off_S + ne * 16 * (1 + KA), h_N.data(), (size_t)KA * ne * 16, cudaMemcpyHostToDevice, st));
  CASQ_CUDA(cudaMemcpyAsync(m->work2.p, m->U.data(), (size_t)d * d * 16, cudaMemcpyHostToDevice, st));
  // the pageable host vectors above are staged by the runtime before cudaMemcpyAsync returns
  DopriGenArgs a;
  memset(&a, 0, sizeof(a));
  a.B = B;
  a.T = TM;
  a.T_out = T_out;
  a.d = d;
  a.KA = KA;
  a.width = width;
  a.rwa = m->rwa;
  a.tpts = (const double*)m->tab_times.p;
  a.keep = (const char*)m->consts.p;
  a.params = params_dev;
  a.y0 = y0_dev;
  a.y0_batched = y0_batched;
  a.out = out_dev;
  a.nsteps = nsteps_dev;
  a.status = status_dev;
  a.n_total = n_total;
  a.dt_sample = m->dt;
  a.inv_dt_sample = 1.0 / m->dt;
  a.rtol = rtol;
  a.atol = atol;
  a.hmax = hmax;
  a.mxstep = mxstep;
  for (int k = 0; k < KA; k++) {
    a.two_pi_nu[k] = 2.0 * M_PI * m->nu[active[k]];
    a.max_abs_freq = std::max(a.max_abs_freq, std::fabs(a.two_pi_nu[k]));
  }
  for (int i = 0; i < d; i++) a.max_abs_freq = std::max(a.max_abs_freq, std::fabs(m->fe[i]));
  a.fe = (const double*)m->work0.p;
  a.ell_col = (const int*)ell;
  a.ell_S = (const double2*)(ell + off_S);
  a.ell_P = a.ell_S + ne;
  a.ell_N = a.ell_P + (size_t)KA * ne;
  a.identityU = m->U_is_identity ? 1 : 0;
  a.U = (const double2*)m->work2.p;
#define DG_LAUNCH(R_, RWA_)                                                                                                   \
  {                                                                                                                           \
    CASQ_CUDA(cudaFuncSetAttribute(sv_dopri5_general_kernel<R_, RWA_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    sv_dopri5_general_kernel<R_, RWA_><<<grid, 32 * DG_WARPS, smem, st>>>(a);                                                 \
  }
  // the batch is walked in windows that bound the envelope table
  const long long Bw = casq_env_table_window(KA, n_total, B);
  if ((rc = m->work3.ensure((size_t)KA * std::max(n_total, 1) * (size_t)Bw * sizeof(double2)))) return rc;
  a.env_tab = (const double2*)m->work3.p;
  for (long long b0 = 0; b0 < B; b0 += Bw) {
    const long long Bc = std::min(Bw, (long long)B - b0);
    CASQ_CUDA(casq_launch_env_table<DG_MAXK>(KA, slots, params_dev, B, b0, Bc, n_total, (double2*)m->work3.p, st));
    a.B = Bc;
    a.y0 = y0_dev + (y0_batched ? b0 * 2 * d : 0);
    a.out = out_dev + b0 * (long long)T_out * 2 * d;
    a.nsteps = nsteps_dev ? nsteps_dev + 2 * b0 : nullptr;
    a.status = status_dev ? status_dev + b0 : nullptr;
    const unsigned grid = (unsigned)((Bc + DG_WARPS - 1) / DG_WARPS);
    if (R == 1 && m->rwa) DG_LAUNCH(1, true)
    else if (R == 1) DG_LAUNCH(1, false)
    else if (R == 2 && m->rwa) DG_LAUNCH(2, true)
    else if (R == 2) DG_LAUNCH(2, false)
    else if (R == 4 && m->rwa) DG_LAUNCH(4, true)
    else if (R == 4) DG_LAUNCH(4, false)
    else if (m->rwa) DG_LAUNCH(8, true)
    else DG_LAUNCH(8, false)
  }
#undef DG_LAUNCH
  CASQ_CUDA(cudaGetLastError());
  return CASQ_OK;
}
