// casq_b200: host-side model preparation behind the C ABI.
//
// Replaces, for the device engine, what qiskit-dynamics' Solver / RotatingFrame /
// rotating_wave_approximation / DynamicsBackend._get_dressed_state_decomposition do when casq calls
// QiskitPulseBackend._get_native_backend (src/casq/backends/qiskit/qiskit_pulse_backend.py:166-189).
// Nothing here is translated from those packages (they are not in /root/reference); the maths is
// SURVEY.md Appendix A.4:   [G_F]_ij = -i e^{i(e_i-e_j)t} [U^+(H(t)-H_f)U]_ij.
#include <cstdarg>
#include <cstring>
#include <algorithm>
#include <numeric>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void casq_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* casq_last_error(void) { return g_err; }
extern "C" int casq_version(void) { return 100; }

int DeviceBuf::ensure(size_t need) {
  if (need <= bytes) return CASQ_OK;
  if (p) cudaFree(p);
  p = nullptr;
  bytes = 0;
  size_t cap = need + need / 4 + 256;
  cudaError_t e = cudaMalloc(&p, cap);
  if (e != cudaSuccess) {
    casq_set_error("cudaMalloc(%zu) failed: %s", cap, cudaGetErrorString(e));
    p = nullptr;
    return CASQ_ERR_NOMEM;
  }
  bytes = cap;
  return CASQ_OK;
}
void DeviceBuf::release() {
  if (p) cudaFree(p);
  p = nullptr;
  bytes = 0;
}

// ---------------------------------------------------------------------------------------------
// Cyclic Jacobi for complex Hermitian matrices.  Each rotation J = D R: D = diag(1, e^{-i phi}) makes
// a_pq real, R is the real Givens rotation that annihilates it.
void casq_eigh(int n, const cplx* Ain, double* evals, cplx* V) {
  std::vector<cplx> A(Ain, Ain + (size_t)n * n);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) V[i * n + j] = (i == j) ? 1.0 : 0.0;
  // symmetrise against input noise
  for (int i = 0; i < n; i++) {
    A[i * n + i] = A[i * n + i].real();
    for (int j = i + 1; j < n; j++) {
      cplx m = 0.5 * (A[i * n + j] + std::conj(A[j * n + i]));
      A[i * n + j] = m;
      A[j * n + i] = std::conj(m);
    }
  }
  double scale = 0;
  for (size_t i = 0; i < A.size(); i++) scale = std::max(scale, std::abs(A[i]));
  if (scale == 0) scale = 1;
  for (int sweep = 0; sweep < 64; sweep++) {
    double off = 0;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) off = std::max(off, std::abs(A[p * n + q]));
    if (off <= 1e-300 || off < 4e-16 * scale * 1e-3) break;
    for (int p = 0; p < n; p++) {
      for (int q = p + 1; q < n; q++) {
        cplx apq = A[p * n + q];
        double mag = std::abs(apq);
        if (mag < 1e-18 * scale) {
          A[p * n + q] = 0;
          A[q * n + p] = 0;
          continue;
        }
        cplx ph = apq / mag;  // e^{i phi}
        double app = A[p * n + p].real(), aqq = A[q * n + q].real();
        double tau = (aqq - app) / (2.0 * mag);
        double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
        double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
        // J columns: Jp = (c, -s e^{-i phi}) on rows (p,q); Jq = (s, c e^{-i phi})
        cplx jpp = c, jqp = -s * std::conj(ph), jpq = s, jqq = c * std::conj(ph);
        for (int k = 0; k < n; k++) {  // A <- A J
          cplx akp = A[k * n + p], akq = A[k * n + q];
          A[k * n + p] = akp * jpp + akq * jqp;
          A[k * n + q] = akp * jpq + akq * jqq;
          cplx vkp = V[k * n + p], vkq = V[k * n + q];
          V[k * n + p] = vkp * jpp + vkq * jqp;
          V[k * n + q] = vkp * jpq + vkq * jqq;
        }
        for (int k = 0; k < n; k++) {  // A <- J^+ A
          cplx apk = A[p * n + k], aqk = A[q * n + k];
          A[p * n + k] = std::conj(jpp) * apk + std::conj(jqp) * aqk;
          A[q * n + k] = std::conj(jpq) * apk + std::conj(jqq) * aqk;
        }
        A[p * n + q] = 0;
        A[q * n + p] = 0;
        A[p * n + p] = A[p * n + p].real();
        A[q * n + q] = A[q * n + q].real();
      }
    }
  }
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](int a, int b) { return A[a * n + a].real() < A[b * n + b].real(); });
  std::vector<cplx> Vs((size_t)n * n);
  for (int c = 0; c < n; c++) {
    evals[c] = A[order[c] * n + order[c]].real();
    for (int r = 0; r < n; r++) Vs[r * n + c] = V[r * n + order[c]];
  }
  std::copy(Vs.begin(), Vs.end(), V);
}

static bool is_diagonal(int n, const cplx* A, double tol) {
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++)
      if (i != j && std::abs(A[i * n + j]) > tol) return false;
  return true;
}

// C = A^+ B A helper pieces
static void matmul(int n, const cplx* A, const cplx* B, cplx* C, bool conjT_A) {
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      cplx acc = 0;
      for (int k = 0; k < n; k++) acc += (conjT_A ? std::conj(A[k * n + i]) : A[i * n + k]) * B[k * n + j];
      C[i * n + j] = acc;
    }
}
static void to_frame_basis(int n, const cplx* U, const cplx* X, cplx* out) {
  std::vector<cplx> tmp((size_t)n * n);
  matmul(n, U, X, tmp.data(), true);     // U^+ X
  matmul(n, tmp.data(), U, out, false);  // (U^+ X) U
}

static int prepare(casq_model* m, const double* frame) {
  const int d = m->d, K = m->K;
  std::vector<cplx> Hf((size_t)d * d, 0.0);
  m->fe.assign(d, 0.0);
  m->U.assign((size_t)d * d, 0.0);
  for (int i = 0; i < d; i++) m->U[i * d + i] = 1.0;
  m->U_is_identity = true;
  if (m->frame_kind == 2) {
    for (int i = 0; i < d; i++) {
      m->fe[i] = frame[i];
      Hf[i * d + i] = frame[i];
    }
  } else if (m->frame_kind == 1) {
    const cplx* F = reinterpret_cast<const cplx*>(frame);
    for (int i = 0; i < d * d; i++) Hf[i] = F[i];
    double sc = 0;
    for (int i = 0; i < d * d; i++) sc = std::max(sc, std::abs(Hf[i]));
    if (is_diagonal(d, Hf.data(), 0.0)) {
      // a diagonal Hermitian frame keeps the bare basis and the bare ordering (eigh would sort ascending;
      // the product only needs *a* diagonalising basis, results are basis-independent)
      for (int i = 0; i < d; i++) m->fe[i] = Hf[i * d + i].real();
    } else {
      casq_eigh(d, Hf.data(), m->fe.data(), m->U.data());
      m->U_is_identity = false;
    }
  }
  // frame-basis operators
  std::vector<cplx> tmp((size_t)d * d);
  for (int i = 0; i < d * d; i++) tmp[i] = m->H0[i] - Hf[i];
  m->S.assign((size_t)d * d, 0.0);
  to_frame_basis(d, m->U.data(), tmp.data(), m->S.data());
  m->P.assign((size_t)K * d * d, 0.0);
  m->N.assign((size_t)K * d * d, 0.0);
  std::vector<cplx> G((size_t)d * d);
  const double two_pi = 2.0 * M_PI;
  for (int k = 0; k < K; k++) {
    to_frame_basis(d, m->U.data(), &m->Hk[(size_t)k * d * d], G.data());
    for (int i = 0; i < d; i++)
      for (int j = 0; j < d; j++) {
        cplx g = 0.5 * G[i * d + j];
        bool keep_p = true, keep_n = true;
        if (m->rwa) {
          double f = (m->fe[i] - m->fe[j]) / two_pi;
          keep_p = std::fabs(m->nu[k] + f) < m->rwa_cutoff;
          keep_n = std::fabs(-m->nu[k] + f) < m->rwa_cutoff;
        }
        m->P[((size_t)k * d + i) * d + j] = keep_p ? g : 0.0;
        m->N[((size_t)k * d + i) * d + j] = keep_n ? g : 0.0;
      }
  }
  if (m->rwa) {
    for (int i = 0; i < d; i++)
      for (int j = 0; j < d; j++) {
        double f = (m->fe[i] - m->fe[j]) / two_pi;
        if (!(std::fabs(f) < m->rwa_cutoff)) m->S[i * d + j] = 0.0;
      }
  }
  // dressed basis of the lab-frame static Hamiltonian, ordered by maximum overlap
  m->de.assign(d, 0.0);
  m->V.assign((size_t)d * d, 0.0);
  if (is_diagonal(d, m->H0.data(), 0.0)) {
    for (int i = 0; i < d; i++) {
      m->de[i] = m->H0[i * d + i].real();
      m->V[i * d + i] = 1.0;
    }
  } else {
    std::vector<double> ev(d);
    std::vector<cplx> W((size_t)d * d);
    casq_eigh(d, m->H0.data(), ev.data(), W.data());
    std::vector<char> seen(d, 0);
    for (int c = 0; c < d; c++) {
      int pos = 0;
      double best = -1;
      for (int r = 0; r < d; r++) {
        double a = std::abs(W[r * d + c]);
        if (a > best) {
          best = a;
          pos = r;
        }
      }
      if (seen[pos]) CASQ_FAIL(CASQ_ERR_INVALID, "dressed-state decomposition: ambiguous overlap at bare state %d", pos);
      seen[pos] = 1;
      cplx phase = std::conj(W[pos * d + c]) / std::abs(W[pos * d + c]);
      for (int r = 0; r < d; r++) m->V[r * d + pos] = W[r * d + c] * phase;
      m->de[pos] = ev[c];
    }
  }
  return CASQ_OK;
}

extern "C" int casq_model_create(casq_model** out, int dim, int n_channels, const double* static_h,
                                 const double* operators, const double* carrier_hz, double dt, int frame_kind,
                                 const double* frame, double rwa_cutoff_hz, int device) {
  if (!out) CASQ_FAIL(CASQ_ERR_INVALID, "casq_model_create: out is NULL");
  *out = nullptr;
  if (dim < 1 || dim > 1024) CASQ_FAIL(CASQ_ERR_INVALID, "casq_model_create: dim %d out of range [1,1024]", dim);
  if (n_channels < 0 || n_channels > 64) CASQ_FAIL(CASQ_ERR_INVALID, "casq_model_create: n_channels %d out of range", n_channels);
  if (!static_h || (n_channels > 0 && (!operators || !carrier_hz))) CASQ_FAIL(CASQ_ERR_INVALID, "casq_model_create: NULL operator input");
  if (!(dt > 0)) CASQ_FAIL(CASQ_ERR_INVALID, "casq_model_create: dt must be > 0");
  if (frame_kind < 0 || frame_kind > 2 || (frame_kind != 0 && !frame)) CASQ_FAIL(CASQ_ERR_INVALID, "casq_model_create: bad rotating frame");
  casq_model* m = new casq_model();
  m->d = dim;
  m->K = n_channels;
  m->dt = dt;
  m->device = device;
  m->frame_kind = frame_kind;
  m->rwa = !std::isnan(rwa_cutoff_hz);
  m->rwa_cutoff = rwa_cutoff_hz;
  const cplx* h0 = reinterpret_cast<const cplx*>(static_h);
  m->H0.assign(h0, h0 + (size_t)dim * dim);
  if (n_channels) {
    const cplx* hk = reinterpret_cast<const cplx*>(operators);
    m->Hk.assign(hk, hk + (size_t)n_channels * dim * dim);
    m->nu.assign(carrier_hz, carrier_hz + n_channels);
  }
  int rc = prepare(m, frame);
  if (rc != CASQ_OK) {
    delete m;
    return rc;
  }
  *out = m;
  return CASQ_OK;
}

extern "C" void casq_model_destroy(casq_model* m) {
  if (!m) return;
  DeviceBuf* bufs[] = {&m->tab, &m->tab_idx, &m->tab_times, &m->pieces_n, &m->pieces_h, &m->consts,
                       &m->work0, &m->work1, &m->work2, &m->work3, &m->slice_buf, &m->ct_chunk};
  for (DeviceBuf* b : bufs) b->release();
  if (m->stage_h) cudaFreeHost(m->stage_h);
  if (m->stage_ev) cudaEventDestroy(m->stage_ev);
  for (int i = 0; i < 3; i++) {
    if (m->aux_stream[i]) cudaStreamDestroy(m->aux_stream[i]);
    if (m->ev_join[i]) cudaEventDestroy(m->ev_join[i]);
  }
  if (m->ev_fork) cudaEventDestroy(m->ev_fork);
  delete m;
}

extern "C" int casq_model_set_dissipators(casq_model* m, int n, const double* ops) {
  if (!m || n < 0 || (n > 0 && !ops)) CASQ_FAIL(CASQ_ERR_INVALID, "casq_model_set_dissipators: bad arguments");
  const int d = m->d;
  m->J = n;
  m->Lfb.assign((size_t)n * d * d, 0.0);
  const cplx* L = reinterpret_cast<const cplx*>(ops);
  for (int j = 0; j < n; j++) to_frame_basis(d, m->U.data(), L + (size_t)j * d * d, &m->Lfb[(size_t)j * d * d]);
  m->tab_key.clear();
  return CASQ_OK;
}

extern "C" int casq_model_dim(const casq_model* m) { return m ? m->d : CASQ_ERR_INVALID; }

extern "C" int casq_model_invalidate_cache(casq_model* m) {
  if (!m) CASQ_FAIL(CASQ_ERR_INVALID, "casq_model_invalidate_cache: NULL model");
  m->tab_key.clear();
  return CASQ_OK;
}

extern "C" int casq_model_get(const casq_model* m, double* frame_evals, double* frame_basis, double* dressed_evals,
                              double* dressed_states) {
  if (!m) CASQ_FAIL(CASQ_ERR_INVALID, "casq_model_get: NULL model");
  const size_t d = m->d;
  if (frame_evals) memcpy(frame_evals, m->fe.data(), d * sizeof(double));
  if (frame_basis) memcpy(frame_basis, m->U.data(), d * d * sizeof(cplx));
  if (dressed_evals) memcpy(dressed_evals, m->de.data(), d * sizeof(double));
  if (dressed_states) memcpy(dressed_states, m->V.data(), d * d * sizeof(cplx));
  return CASQ_OK;
}

// ---------------------------------------------------------------------------------------------
// Python/numpy float floor-division (a // b), the operation DiscreteSignal uses to pick a sample.
long long casq_py_floor_divide(double a, double b) {
  double mod = std::fmod(a, b);
  double div = (a - mod) / b;
  if (mod != 0.0 && ((b < 0) != (mod < 0))) div -= 1.0;
  double fl;
  if (div != 0.0) {
    fl = std::floor(div);
    if (div - fl > 0.5) fl += 1.0;
  } else {
    fl = 0.0;
  }
  return (long long)fl;
}

// The same quotient without the iterative fmod: a guess from one multiplication, corrected by the EXACT sign of the remainder
// a - q b (one FMA; the device steppers use the same rule, dopri5_common.cuh).  Only a remainder that rounds to exactly b is
// ambiguous: that case takes the reference path above.
static inline long long floor_divide_fast(double a, double b, double inv_b) {
  double q = std::floor(a * inv_b);
  const double r = std::fma(-q, b, a);
  if (r == b || !(b > 0)) return casq_py_floor_divide(a, b);
  if (r < 0.0)
    q -= 1.0;
  else if (r > b)
    q += 1.0;
  return (long long)q;
}

int casq_build_time_grid(double t0, double t1, double max_dt, int T, const double* t_eval, double dt,
                         int n_total, TimeGrid* g) {
  if (!(max_dt > 0)) CASQ_FAIL(CASQ_ERR_INVALID, "fixed-step solve needs max_dt > 0");
  std::vector<double> pts;
  g->out_keep.clear();
  if (!t_eval) {
    if (T != 2) CASQ_FAIL(CASQ_ERR_INVALID, "t_eval is NULL: T must be 2 (t0 and t1), got %d", T);
    pts = {t0, t1};
    g->out_keep = {1, 1};
  } else {
    if (T < 1) CASQ_FAIL(CASQ_ERR_INVALID, "t_eval needs at least one point");
    for (int i = 1; i < T; i++)
      if (!(t_eval[i] > t_eval[i - 1])) CASQ_FAIL(CASQ_ERR_INVALID, "t_eval must be strictly increasing");
    if (t_eval[0] < t0 || t_eval[T - 1] > t1) CASQ_FAIL(CASQ_ERR_INVALID, "t_eval must lie inside [t0, t1]");
    if (t_eval[0] != t0) {
      pts.push_back(t0);
      g->out_keep.push_back(0);
    }
    for (int i = 0; i < T; i++) {
      pts.push_back(t_eval[i]);
      g->out_keep.push_back(1);
    }
    if (t_eval[T - 1] != t1) {
      pts.push_back(t1);
      g->out_keep.push_back(0);
    }
  }
  g->out_times = pts;
  const int np = (int)pts.size() - 1;
  g->piece_nsteps.assign(np, 0);
  g->piece_h.assign(np, 0.0);
  g->times.clear();
  g->sample_idx.clear();
  {
    size_t total = 0;
    for (int p = 0; p < np; p++) total += 2 * (size_t)std::min(1e9, std::fabs((pts[p + 1] - pts[p]) / max_dt) + 2.0) + 1;
    g->times.reserve(total);
  }
  for (int p = 0; p < np; p++) {
    double delta = pts[p + 1] - pts[p];
    long long n = (long long)std::fabs(delta / max_dt);
    if (n == 0)
      n = 1;
    else if (std::fabs(delta / (double)n) / max_dt > 1.0 + 1e-15)
      n += 1;
    if (n > (1LL << 30)) CASQ_FAIL(CASQ_ERR_INVALID, "too many steps in one piece (%lld)", n);
    double h = delta / (double)n;
    g->piece_nsteps[p] = (int)n;
    g->piece_h[p] = h;
    double t = pts[p];
    double h2 = 0.5 * h;
    for (long long s = 0; s < n; s++) {
      g->times.push_back(t);
      g->times.push_back(t + h2);
      t = t + h;
    }
    g->times.push_back(t);
  }
  g->sample_idx.resize(g->times.size());
  const double inv_dt = 1.0 / dt;
  for (size_t i = 0; i < g->times.size(); i++) {
    long long q = floor_divide_fast(g->times[i], dt, inv_dt);
    if (q < -1) q = -1;
    if (q > n_total) q = n_total;
    g->sample_idx[i] = (int)q;
  }
  return CASQ_OK;
}
