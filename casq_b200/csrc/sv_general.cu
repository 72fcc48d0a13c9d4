// casq_b200: fixed-step RK4 state-vector stepper for GENERAL Hilbert-space dimension (5..64), one trajectory
// per lane group (16 lanes for d<=16, 32 lanes x ceil(d/32) rows otherwise), sparse-row matvec through shared
// memory.
//
// Same role as sv_small.cu (QiskitPulseBackend.solve with ODESolverMethod.QISKIT_DYNAMICS_RK4,
// src/casq/backends/qiskit/qiskit_pulse_backend.py:158-162) for coupled-transmon models, e.g. the 2-transmon
// d=9 model of BASELINE cfg 3 integrated column by column, or 3 transmons (d=27).
//
// Per stage, lane i (row i):   z'_i = conj(p_i) v_i ;  w_i = sum_q A_iq z'_{col(i,q)} ;  k_i = -i p_i w_i
// with p_i = e^{i e_i t} from the shared stage table (so the d^2 Bohr phases cost 2 complex multiplies per row)
// and A = S + sum_k (z_k P_k + conj(z_k) N_k) assembled on the fly from ELL-packed rows of the frame-basis
// operators (X-type drive and exchange-coupling operators have <= 4 non-zeros per row).
#include <algorithm>
#include <cstring>
#include <string>

#include "common.cuh"
#include "envelope.cuh"

#define GEN_MAXK 4

struct GenArgs {
  const double* table;  // [NT][W], W = 2*KA + 2*d : carriers then p_i
  const int* idx;
  const int* piece_n;
  const double* piece_h;
  const char* keep;
  int n_pieces;
  long long B;
  int T, d, KA, width, W;
  const double* params;
  const double* y0;
  int y0_batched;
  double* out;
  DevSlots slots;
  // ELL tables in global memory (copied to shared memory by each block): [width][d]
  const int* ell_col;
  const double2* ell_S;  // [width][d]
  const double2* ell_P;  // [KA][width][d]
  const double2* ell_N;  // [KA][width][d]   (RWA only)
  int identityU;
  const double2* U;  // [d][d] frame basis (row-major), only when !identityU
};

template <int G, int R, bool RWA>
__global__ void __launch_bounds__(128) sv_rk4_general_kernel(const GenArgs a) {
  constexpr int GROUPS = 128 / G;
  const int d = a.d, KA = a.KA, width = a.width;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: col[width*d] int | S[width*d] | P[KA*width*d] | (N[KA*width*d]) | zbuf[GROUPS][R*G] | ybuf (for U transforms)
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

  const int lane = threadIdx.x % G, grp = threadIdx.x / G;
  const long long b_raw = (long long)blockIdx.x * GROUPS + grp;
  const bool live = b_raw < a.B;
  const long long b = live ? b_raw : a.B - 1;
  double2* zb = s_z + (size_t)grp * (R * G);
  const unsigned gmask = (G == 32) ? 0xffffffffu : (0xffffu << (16 * ((threadIdx.x % 32) / 16)));

  // rows owned by this lane: i = lane + r*G
  double yr[R], yi[R];
  bool own[R];
#pragma unroll
  for (int r = 0; r < R; r++) own[r] = (lane + r * G) < d;

  // y <- U^+ y0
  {
    const double* src = a.y0 + (a.y0_batched ? b * 2 * d : 0);
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int i = lane + r * G;
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
  double* outp = a.out + b * (long long)a.T * 2 * d;
  int out_slot = 0;
  auto emit = [&](void) {
    double* o = outp + (long long)out_slot * 2 * d;
    if (a.identityU) {
#pragma unroll
      for (int r = 0; r < R; r++)
        if (own[r] && live) {
          o[2 * (lane + r * G)] = yr[r];
          o[2 * (lane + r * G) + 1] = yi[r];
        }
    } else {
      __syncwarp(gmask);
#pragma unroll
      for (int r = 0; r < R; r++)
        if (own[r]) zb[lane + r * G] = make_double2(yr[r], yi[r]);
      __syncwarp(gmask);
#pragma unroll
      for (int r = 0; r < R; r++) {
        const int i = lane + r * G;
        if (own[r] && live) {
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
      __syncwarp(gmask);
    }
    out_slot++;
  };
  if (a.keep[0]) emit();

  int cur_idx = -2;
  dcplx env[GEN_MAXK];
#pragma unroll
  for (int k = 0; k < GEN_MAXK; k++) env[k].re = env[k].im = 0.0;
  auto env_at = [&](int gidx) {
    if (gidx != cur_idx) {
      cur_idx = gidx;
      casq_channel_envelopes<GEN_MAXK>(a.slots, a.params, a.B, b, gidx, env);
    }
  };

  // k = -i p (A (conj(p) v))
  auto rhs = [&](long long pos, const double* vr, const double* vi, double* kr, double* ki) {
    const double* ent = a.table + pos * a.W;
    double zr[GEN_MAXK], zi[GEN_MAXK];
#pragma unroll
    for (int k = 0; k < GEN_MAXK; k++) {
      if (k < KA) {
        const double c = __ldg(ent + 2 * k), s = __ldg(ent + 2 * k + 1);
        zr[k] = env[k].re * c - env[k].im * s;
        zi[k] = env[k].re * s + env[k].im * c;
      } else {
        zr[k] = zi[k] = 0.0;
      }
    }
    double pr[R], pi[R];
    __syncwarp(gmask);
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int i = lane + r * G;
      pr[r] = 1.0;
      pi[r] = 0.0;
      if (own[r]) {
        const double2 p = __ldg(reinterpret_cast<const double2*>(ent + 2 * KA) + i);
        pr[r] = p.x;
        pi[r] = p.y;
        // conj(p) * v
        zb[i] = make_double2(p.x * vr[r] + p.y * vi[r], p.x * vi[r] - p.y * vr[r]);
      }
    }
    __syncwarp(gmask);
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int i = lane + r * G;
      double wr = 0.0, wi = 0.0;
      if (own[r]) {
        for (int q = 0; q < width; q++) {
          const int e = q * d + i;
          const int col = s_col[e];
          double2 A = s_S[e];
#pragma unroll
          for (int k = 0; k < GEN_MAXK; k++) {
            if (k < KA) {
              const double2 P = s_P[(size_t)k * width * d + e];
              if (RWA) {
                const double2 N = s_N[(size_t)k * width * d + e];
                A.x += zr[k] * (P.x + N.x) - zi[k] * (P.y - N.y);
                A.y += zr[k] * (P.y + N.y) + zi[k] * (P.x - N.x);
              } else {  // P holds G (= 2 * G/2): real signal
                A.x = fma(zr[k], P.x, A.x);
                A.y = fma(zr[k], P.y, A.y);
              }
            }
          }
          const double2 z = zb[col];
          wr += A.x * z.x - A.y * z.y;
          wi += A.x * z.y + A.y * z.x;
        }
      }
      // k = -i p w
      const double tr = pr[r] * wr - pi[r] * wi, ti = pr[r] * wi + pi[r] * wr;
      kr[r] = ti;
      ki[r] = -tr;
    }
  };

  long long pos = 0;
  for (int p = 0; p < a.n_pieces; p++) {
    const int n = a.piece_n[p];
    const double h = a.piece_h[p];
    const double h2 = 0.5 * h, h6 = h * (1.0 / 6.0), h3 = h * (1.0 / 3.0);
    for (int s = 0; s < n; s++) {
      const int ia = __ldg(a.idx + pos), ib = __ldg(a.idx + pos + 1), ic = __ldg(a.idx + pos + 2);
      double kr[R], ki[R], ar[R], ai[R], tr[R], ti[R];
      env_at(ia);
      rhs(pos, yr, yi, kr, ki);
#pragma unroll
      for (int r = 0; r < R; r++) {
        ar[r] = fma(h6, kr[r], yr[r]);
        ai[r] = fma(h6, ki[r], yi[r]);
        tr[r] = fma(h2, kr[r], yr[r]);
        ti[r] = fma(h2, ki[r], yi[r]);
      }
      env_at(ib);
      rhs(pos + 1, tr, ti, kr, ki);
#pragma unroll
      for (int r = 0; r < R; r++) {
        ar[r] = fma(h3, kr[r], ar[r]);
        ai[r] = fma(h3, ki[r], ai[r]);
        tr[r] = fma(h2, kr[r], yr[r]);
        ti[r] = fma(h2, ki[r], yi[r]);
      }
      rhs(pos + 1, tr, ti, kr, ki);
#pragma unroll
      for (int r = 0; r < R; r++) {
        ar[r] = fma(h3, kr[r], ar[r]);
        ai[r] = fma(h3, ki[r], ai[r]);
        tr[r] = fma(h, kr[r], yr[r]);
        ti[r] = fma(h, ki[r], yi[r]);
      }
      env_at(ic);
      rhs(pos + 2, tr, ti, kr, ki);
#pragma unroll
      for (int r = 0; r < R; r++) {
        yr[r] = fma(h6, kr[r], ar[r]);
        yi[r] = fma(h6, ki[r], ai[r]);
      }
      pos += 2;
    }
    pos += 1;
    if (a.keep[p + 1]) emit();
  }
}

// table: per time point [KA x (cos, sin)] [d x p_i = e^{i e_i t}]
__global__ void sv_general_table_kernel(const double* __restrict__ times, long long NT, int KA, int d, int W,
                                        const double* __restrict__ two_pi_nu, const double* __restrict__ fe,
                                        double* __restrict__ out) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int per = KA + d;
  if (g >= NT * per) return;
  const long long i = g / per;
  const int j = (int)(g - i * per);
  const double t = times[i];
  const double arg = (j < KA ? two_pi_nu[j] : fe[j - KA]) * t;
  double s, c;
  sincos(arg, &s, &c);
  out[i * W + 2 * j] = c;
  out[i * W + 2 * j + 1] = s;
}

bool casq_general_supported(const casq_model* m, int n_active) {
  return m->d >= 2 && m->d <= 256 && n_active >= 1 && n_active <= GEN_MAXK;
}

int casq_general_rk4(casq_model* m, int KA, const int* active, const DevSlots& slots, int n_total, long long B,
                     const double* params_dev, const double* y0_dev, int y0_batched, double t0, double t1, double max_dt,
                     int T, const double* t_eval, double* out_dev, cudaStream_t st) {
  const int d = m->d;
  // ---- ELL pack of the union sparsity pattern
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
  const int G = d <= 16 ? 16 : 32;
  int R = (d + G - 1) / G;
  if (R > 4) R = 8;  // instantiated row counts: 1, 2, 3, 4, 8 (dim <= 256: the full 5-transmon model is 243)
  const int groups = 128 / G;
  size_t smem = (((size_t)ne * sizeof(int) + 15) & ~(size_t)15) + ne * 16 * (1 + (size_t)KA * (m->rwa ? 2 : 1)) +
                (size_t)groups * R * G * 16;
  if (smem > 200 * 1024)
    CASQ_FAIL(CASQ_ERR_UNSUPPORTED, "general RK4: operators too dense for shared memory (dim %d, row width %d, %d channels): large models need a diagonal rotating frame (rotating_frame = diag(static_operator), casq's from_backend default)", d, width, KA);

  // ---- time grid + table
  TimeGrid g;
  int rc = casq_build_time_grid(t0, t1, max_dt, T, t_eval, m->dt, n_total, &g);
  if (rc != CASQ_OK) return rc;
  const long long NT = (long long)g.times.size();
  const int n_pieces = (int)g.piece_nsteps.size();
  const int W = 2 * KA + 2 * d;
  std::vector<char> keep(g.out_keep.begin(), g.out_keep.end());
  m->tab_key.clear();  // this path rebuilds its table every call (it shares buffers with the small path)
  std::vector<double> consts(KA + d);
  for (int k = 0; k < KA; k++) consts[k] = 2.0 * M_PI * m->nu[active[k]];
  for (int i = 0; i < d; i++) consts[KA + i] = m->fe[i];
  const size_t ell_bytes = ne * sizeof(int) + 16 + ne * 16 * (1 + 2 * (size_t)KA);
  if ((rc = m->tab_times.ensure(NT * sizeof(double)))) return rc;
  if ((rc = m->tab_idx.ensure(NT * sizeof(int)))) return rc;
  if ((rc = m->tab.ensure((size_t)NT * W * sizeof(double)))) return rc;
  if ((rc = m->pieces_n.ensure(n_pieces * sizeof(int)))) return rc;
  if ((rc = m->pieces_h.ensure(n_pieces * sizeof(double)))) return rc;
  if ((rc = m->consts.ensure(keep.size()))) return rc;
  if ((rc = m->work0.ensure(consts.size() * sizeof(double)))) return rc;
  if ((rc = m->work1.ensure(ell_bytes))) return rc;
  if ((rc = m->work2.ensure((size_t)d * d * 16))) return rc;
  CASQ_CUDA(cudaMemcpyAsync(m->tab_times.p, g.times.data(), NT * sizeof(double), cudaMemcpyHostToDevice, st));
  CASQ_CUDA(cudaMemcpyAsync(m->tab_idx.p, g.sample_idx.data(), NT * sizeof(int), cudaMemcpyHostToDevice, st));
  CASQ_CUDA(cudaMemcpyAsync(m->pieces_n.p, g.piece_nsteps.data(), n_pieces * sizeof(int), cudaMemcpyHostToDevice, st));
  CASQ_CUDA(cudaMemcpyAsync(m->pieces_h.p, g.piece_h.data(), n_pieces * sizeof(double), cudaMemcpyHostToDevice, st));
  CASQ_CUDA(cudaMemcpyAsync(m->consts.p, keep.data(), keep.size(), cudaMemcpyHostToDevice, st));
  CASQ_CUDA(cudaMemcpyAsync(m->work0.p, consts.data(), consts.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  char* ell = (char*)m->work1.p;
  const size_t off_S = (ne * sizeof(int) + 15) & ~(size_t)15;
  CASQ_CUDA(cudaMemcpyAsync(ell, h_col.data(), ne * sizeof(int), cudaMemcpyHostToDevice, st));
  CASQ_CUDA(cudaMemcpyAsync(ell + off_S, h_S.data(), ne * 16, cudaMemcpyHostToDevice, st));
  CASQ_CUDA(cudaMemcpyAsync(ell + off_S + ne * 16, h_P.data(), (size_t)KA * ne * 16, cudaMemcpyHostToDevice, st));
  CASQ_CUDA(cudaMemcpyAsync(ell + off_S + ne * 16 * (1 + KA), h_N.data(), (size_t)KA * ne * 16, cudaMemcpyHostToDevice, st));
  CASQ_CUDA(cudaMemcpyAsync(m->work2.p, m->U.data(), (size_t)d * d * 16, cudaMemcpyHostToDevice, st));
  {
    const long long nthreads = NT * (KA + d);
    sv_general_table_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>(
        (const double*)m->tab_times.p, NT, KA, d, W, (const double*)m->work0.p, (const double*)m->work0.p + KA,
        (double*)m->tab.p);
    CASQ_CUDA(cudaGetLastError());
  }
  GenArgs a;
  memset(&a, 0, sizeof(a));
  a.table = (const double*)m->tab.p;
  a.idx = (const int*)m->tab_idx.p;
  a.piece_n = (const int*)m->pieces_n.p;
  a.piece_h = (const double*)m->pieces_h.p;
  a.keep = (const char*)m->consts.p;
  a.n_pieces = n_pieces;
  a.B = B;
  a.T = T;
  a.d = d;
  a.KA = KA;
  a.width = width;
  a.W = W;
  a.params = params_dev;
  a.y0 = y0_dev;
  a.y0_batched = y0_batched;
  a.out = out_dev;
  a.slots = slots;
  a.ell_col = (const int*)ell;
  a.ell_S = (const double2*)(ell + off_S);
  a.ell_P = a.ell_S + ne;
  a.ell_N = a.ell_P + (size_t)KA * ne;
  a.identityU = m->U_is_identity ? 1 : 0;
  a.U = (const double2*)m->work2.p;
  const unsigned grid = (unsigned)((B + groups - 1) / groups);
#define GEN_LAUNCH(G_, R_)                                                                                              \
  if (G == G_ && R == R_) {                                                                                             \
    if (m->rwa) {                                                                                                       \
      CASQ_CUDA(cudaFuncSetAttribute(sv_rk4_general_kernel<G_, R_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      sv_rk4_general_kernel<G_, R_, true><<<grid, 128, smem, st>>>(a);                                                  \
    } else {                                                                                                            \
      CASQ_CUDA(cudaFuncSetAttribute(sv_rk4_general_kernel<G_, R_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      sv_rk4_general_kernel<G_, R_, false><<<grid, 128, smem, st>>>(a);                                                 \
    }                                                                                                                   \
  }
  GEN_LAUNCH(16, 1) GEN_LAUNCH(32, 1) GEN_LAUNCH(32, 2) GEN_LAUNCH(32, 3) GEN_LAUNCH(32, 4) GEN_LAUNCH(32, 8)
#undef GEN_LAUNCH
  CASQ_CUDA(cudaGetLastError());
  return CASQ_OK;
}
