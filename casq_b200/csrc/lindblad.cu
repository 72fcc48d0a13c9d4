// casq_b200: Lindblad master equation, fixed-step RK4 on batches of density matrices, structured
// ("shift-class") application of the frame-rotated superoperator.  Built for BASELINE cfg 4: the 5-transmon
// ibmq_manila model (dim 243) with T1/T2 dissipators, where rho (0.945 MB) lives in HBM and a dense
// superoperator (55.8 GB) or dense commutators (2.3e8 flop per stage) are out of the question.
//
// The reference never builds a noise model (no dissipators are passed to Solver,
// src/casq/backends/qiskit/qiskit_pulse_backend.py:174-183; NoiseModel is a box in
// docs/source/_static/images/casq-object-model.gv:9), so this is the additive API north_star item (4) asks for;
// the maths is qiskit-dynamics' LindbladModel in a diagonal rotating frame (last bullet):
//     rho' = -i[H_F(t), rho] + sum_j ( L_j(t) rho L_j(t)^+ - 1/2 { L_j(t)^+ L_j(t), rho } ),
// every operator element carrying its Bohr phase e^{i(e_a - e_b)t}.
//
// Structure used (frame basis = bare tensor-product basis, i.e. a diagonal frame -- casq's from_backend default,
// SURVEY §5 quirk 1):
//   * one-sided part: with alpha(a,b,t) = -i H_F[a,b](t) - 1/2 D_F[a,b](t), D = sum_j L_j^+ L_j,
//         rho'[r,c] += sum_b alpha(r,b) rho[b,c] + sum_b conj(alpha(c,b)) rho[r,b].
//     Elements are grouped into SHIFT CLASSES by delta = b - a (local operators on a product basis have a handful
//     of distinct deltas: +-3^q for a drive on qubit q, +-(3^q - 3^(q+1)) for exchange coupling, 0 for diagonals);
//     within a class a row has at most one element, whose value and "time function" (signal kind x Bohr frequency)
//     are table look-ups.  One class = one gathered rho element per side.
//   * two-sided part: L_j rho L_j^+ [r,c] = sum over class pairs (q1,q2) of L_j of u_q1(r) conj(u_q2(c))
//     rho[r+delta1, c+delta2].
// Per RK4 stage ONE fused kernel on 64x64 tile pairs (k = M + M^+ with M = A rho + two-sided terms, plus the RK4
// combination with four rolling buffers, lindblad_stage.cuh); per step one coefficient kernel (lindblad_kernels.cuh).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>

#include "lindblad_kernels.cuh"
#include "lindblad_stage.cuh"
#include "lindblad_chain.cuh"

__global__ void sv_general_table_kernel(const double* __restrict__ times, long long NT, int KA, int d, int W,
                                        const double* __restrict__ two_pi_nu, const double* __restrict__ fe,
                                        double* __restrict__ out);
int casq_prepare_slots(const casq_model* m, int n_slots, const casq_pulse_slot* slots, DevSlots* ds, int* active,
                       int* n_active, int* n_total);

// RK4 stages that run the TMA-staged kernel by default (bit 0 = stage 1); see lindblad_run_chain
#ifndef LC_DEFAULT_TMA_STAGES
#define LC_DEFAULT_TMA_STAGES 15
#endif

namespace {
struct Atom {
  int row, col, sig;
  cplx val;
  double omega;
};
struct ClassTab {
  std::vector<int> delta;
  std::vector<short> code;  // [nclass][d]
  std::vector<cplx> val;    // [nclass][d]
};
struct Builder {
  int d;
  std::vector<double> freqs;                   // distinct Bohr frequencies
  std::vector<std::pair<int, int>> tfs;        // (sig, freq id)
  std::map<std::pair<int, int>, int> tf_index;
  double tol;
  int freq_id(double w) {
    if (w == 0.0) return -1;
    for (size_t i = 0; i < freqs.size(); i++)
      if (std::fabs(freqs[i] - w) <= tol) return (int)i;
    freqs.push_back(w);
    return (int)freqs.size() - 1;
  }
  int tf_id(int sig, double w) {
    std::pair<int, int> key(sig, freq_id(w));
    auto it = tf_index.find(key);
    if (it != tf_index.end()) return it->second;
    tfs.push_back(key);
    tf_index[key] = (int)tfs.size() - 1;
    return (int)tfs.size() - 1;
  }
  // place atoms into shift classes: key delta, first class whose row is free
  void place(const std::vector<Atom>& atoms, ClassTab* tab) {
    for (const Atom& at : atoms) {
      const int delta = at.col - at.row;
      int q = -1;
      for (size_t c = 0; c < tab->delta.size(); c++)
        if (tab->delta[c] == delta && tab->code[c * d + at.row] < 0) {
          q = (int)c;
          break;
        }
      if (q < 0) {
        q = (int)tab->delta.size();
        tab->delta.push_back(delta);
        tab->code.resize((size_t)(q + 1) * d, (short)-1);
        tab->val.resize((size_t)(q + 1) * d, cplx(0.0));
      }
      tab->code[(size_t)q * d + at.row] = (short)tf_id(at.sig, at.omega);
      tab->val[(size_t)q * d + at.row] = at.val;
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------
// Chain-of-three-level-sites detection (lindblad_chain.cuh).  Every element of A = -iH_F - D_F/2 and of the dissipators must
// be (a) diagonal and time-independent, (b) a single-site ladder element x_q -> x_q +- 1, or (c) a nearest-neighbour
// exchange element (x_q, x_q+1) -> (x_q -+ 1, x_q+1 +- 1), with value and time function depending only on the digits of the
// sites it touches; each dissipator is either diagonal or a single-site ladder operator in one direction.
struct ChainPlan {
  bool ok = false;
  int nq = 0;
  unsigned drv_mask = 0;
  int n2 = 0, two_q[LC_MAX2], two_dir[LC_MAX2];
  int NS = 0, NC = 0, u_base = 0, w_base = 0;
  std::vector<int> slot_start;
  std::vector<cplx> atom_val;
  std::vector<int> atom_code;
  std::vector<cplx> G;  // [d*ld]
};

struct SlotRows {  // what each row contributes to one slot: (time-function code -> value)
  std::map<int, std::map<int, cplx>> rows;
  long long expected = 0;
};

static bool lc_classify(int nq, int r, int c, int* slot) {
  // returns false when (r, c) is neither a single-site ladder element nor a nearest-neighbour exchange element
  int qs[2], df[2], nd = 0, xr[LC_MAXQ];
  int rr = r, cc = c;
  for (int q = 0; q < nq; q++) {
    const int a = rr % 3, b = cc % 3;
    xr[q] = a;
    rr /= 3;
    cc /= 3;
    if (a != b) {
      if (nd == 2 || std::abs(b - a) != 1) return false;
      qs[nd] = q;
      df[nd] = b - a;
      nd++;
    }
  }
  if (nd == 1) {
    const int q = qs[0];
    *slot = df[0] > 0 ? (q * 2 + 0) * 2 + xr[q] : (q * 2 + 1) * 2 + xr[q] - 1;
    return true;
  }
  if (nd == 2 && qs[1] == qs[0] + 1 && df[0] == -df[1]) {
    const int q = qs[0];
    const int jbase = 4 * nq;
    if (df[0] < 0)  // c = r + 3^(q+1) - 3^q: n = x_q(r) - 1, m = x_(q+1)(r)
      *slot = jbase + ((q * 2 + 0) * 2 + xr[q] - 1) * 2 + xr[q + 1];
    else            // c = r - 3^(q+1) + 3^q: n = x_q(r), m = x_(q+1)(r) - 1
      *slot = jbase + ((q * 2 + 1) * 2 + xr[q]) * 2 + xr[q + 1] - 1;
    return true;
  }
  return false;
}

static bool lc_consistent(const SlotRows& sr, double tol, std::map<int, cplx>* rep) {
  if (sr.rows.empty()) {
    rep->clear();
    return true;
  }
  if ((long long)sr.rows.size() != sr.expected) return false;  // some row that owns this slot has no element
  const std::map<int, cplx>& first = sr.rows.begin()->second;
  for (const auto& kv : sr.rows) {
    if (kv.second.size() != first.size()) return false;
    auto it = first.begin();
    for (const auto& e : kv.second) {
      if (e.first != it->first || std::abs(e.second - it->second) > tol * (1.0 + std::abs(it->second))) return false;
      ++it;
    }
  }
  *rep = first;
  return true;
}

static ChainPlan lc_analyze(const casq_model* m, Builder& bld, const std::vector<Atom>& one) {
  ChainPlan P;
  const int d = m->d;
  int nq = 0, dd = 1;
  while (dd < d) {
    dd *= 3;
    nq++;
  }
  if (dd != d || nq < 3 || nq > LC_MAXQ) return P;
  if (getenv("CASQ_LB_GENERIC")) return P;  // test / tuning knob: force the generic shift-class kernel
  P.nq = nq;
  const int NX = 4 * nq, NJ = 8 * (nq - 1);
  std::vector<SlotRows> slots(NX + NJ);
  for (int s = 0; s < NX; s++) slots[s].expected = d / 3;
  for (int s = 0; s < NJ; s++) slots[NX + s].expected = d / 9;
  std::vector<cplx> adiag(d, cplx(0.0));
  for (const Atom& at : one) {
    const int code = bld.tf_id(at.sig, at.omega);
    if (at.row == at.col) {
      if (bld.tfs[code].first != 0 || bld.tfs[code].second != -1) return P;  // time-dependent diagonal element
      adiag[at.row] += at.val;
      continue;
    }
    int slot;
    if (!lc_classify(nq, at.row, at.col, &slot)) return P;
    slots[slot].rows[at.row][code] += at.val;
  }
  const double tol = 1e-12;
  std::vector<std::map<int, cplx>> reps(NX + NJ);
  for (int s = 0; s < NX + NJ; s++)
    if (!lc_consistent(slots[s], tol, &reps[s])) return P;
  for (int q = 0; q < nq; q++)
    for (int k = 0; k < 4; k++)
      if (!reps[q * 4 + k].empty()) P.drv_mask |= 1u << q;
  // dissipators: diagonal ones go into G, single-site ladder ones become two-sided terms
  std::vector<std::vector<cplx>> diagL;
  std::vector<std::map<int, cplx>> ureps;
  for (int j = 0; j < m->J; j++) {
    const cplx* L = &m->Lfb[(size_t)j * d * d];
    bool any_diag = false, any_off = false;
    int q_ = -1, dir_ = 0;
    SlotRows u[2];
    u[0].expected = u[1].expected = d / 3;
    for (int a_ = 0; a_ < d; a_++)
      for (int b_ = 0; b_ < d; b_++) {
        const cplx v = L[a_ * d + b_];
        if (std::abs(v) == 0.0) continue;
        if (a_ == b_) {
          any_diag = true;
          continue;
        }
        any_off = true;
        int slot;
        if (!lc_classify(nq, a_, b_, &slot) || slot >= NX) return P;
        const int q = slot / 4, dirbit = (slot / 2) & 1, n = slot & 1;
        const int dir = dirbit == 0 ? +1 : -1;
        if (dir < 0) return P;  // raising-type dissipators (thermal excitation) run the generic kernel
        if (q_ < 0) {
          q_ = q;
          dir_ = dir;
        } else if (q_ != q || dir_ != dir) {
          return P;
        }
        const int code = bld.tf_id(0, m->fe[a_] - m->fe[b_]);
        u[n].rows[a_][code] += v;
      }
    if (any_diag && any_off) return P;
    if (any_off) {
      if (P.n2 >= LC_MAX2) return P;
      std::map<int, cplx> r0, r1;
      if (!lc_consistent(u[0], tol, &r0) || !lc_consistent(u[1], tol, &r1)) return P;
      P.two_q[P.n2] = q_;
      P.two_dir[P.n2] = dir_;
      P.n2++;
      ureps.push_back(r0);
      ureps.push_back(r1);
    } else if (any_diag) {
      std::vector<cplx> l(d);
      for (int a_ = 0; a_ < d; a_++) l[a_] = L[a_ * d + a_];
      diagL.push_back(l);
    }
  }
  // flatten the slots
  P.u_base = NX + NJ;
  P.NS = P.u_base + 2 * P.n2;
  P.w_base = P.NS;
  P.NC = P.NS + 4 * P.n2;
  P.slot_start.push_back(0);
  auto put = [&](const std::map<int, cplx>& rep) {
    for (const auto& e : rep) {
      P.atom_code.push_back(e.first);
      P.atom_val.push_back(e.second);
    }
    P.slot_start.push_back((int)P.atom_code.size());
  };
  for (int s = 0; s < NX + NJ; s++) put(reps[s]);
  for (const auto& r : ureps) put(r);
  // G[r,c] = A_rr + conj(A_cc) + sum over diagonal L_j of l_j(r) conj(l_j(c))
  const int ld = (d / LC_T) * LC_TS;  // the chain path's tile-major storage (lb_idx)
  P.G.assign((size_t)d * ld, cplx(0.0));
  for (int r = 0; r < d; r++)
    for (int c = 0; c < d; c++) {
      cplx g = adiag[r] + std::conj(adiag[c]);
      for (const auto& l : diagL) g += l[r] * std::conj(l[c]);
      P.G[((size_t)(r / LC_T) * (d / LC_T) + c / LC_T) * (LC_T * LC_TS) + (r % LC_T) * LC_TS + c % LC_T] = g;  // tile-major (lb_idx)
    }
  P.ok = true;
  return P;
}

// The chain path: same buffers, time grid, stream parts and outputs as the generic path below; per RK4 step one small
// coefficient kernel and four warp-per-tile-pair stage kernels (lindblad_chain.cuh).
static int lindblad_run_chain(casq_model* m, const ChainPlan& cp, Builder& bld, int KA, const int* active, const DevSlots& ds,
                              int n_total, int64_t B, const double* params_dev, const double* rho0_dev, int rho0_batched,
                              double t0, double t1, double max_dt, int T, const double* t_eval, double* out_rho_dev,
                              double* out_pops_dev, cudaStream_t st) {
  const int d = m->d, ld = (d / LC_T) * LC_TS;  // tile-major storage (lb_idx): d * ld elements per matrix
  const int tw = LC_T, tpad = LC_TS - LC_T;
  int rc;
  TimeGrid g;
  rc = casq_build_time_grid(t0, t1, max_dt, T, t_eval, m->dt, n_total, &g);
  if (rc != CASQ_OK) return rc;
  const long long NT = (long long)g.times.size();
  const int NF = (int)bld.freqs.size(), NTF = (int)bld.tfs.size();
  const int W = 2 * KA + 2 * NF;
  m->tab_key.clear();
  std::vector<double> consts(KA + std::max(NF, 1));
  for (int k = 0; k < KA; k++) consts[k] = 2.0 * M_PI * m->nu[active[k]];
  for (int f = 0; f < NF; f++) consts[KA + f] = bld.freqs[f];
  std::vector<int> tf_sig(NTF), tf_freq(NTF);
  for (int n = 0; n < NTF; n++) {
    tf_sig[n] = bld.tfs[n].first;
    tf_freq[n] = bld.tfs[n].second;
  }
  const int ntile = d / LC_T;
  std::vector<int> pair_ij;
  for (int I = 0; I < ntile; I++)  // pairs of one trajectory are adjacent tasks: its tiles stay hot in L2 / L1
    for (int J = I; J < ntile; J++) {
      pair_ij.push_back(I);
      pair_ij.push_back(J);
    }
  const int npairs = (int)pair_ij.size() / 2;
  auto align16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
  size_t off = 0;
  const size_t o_tfsig = off; off = align16(off + std::max(NTF, 1) * sizeof(int));
  const size_t o_tffreq = off; off = align16(off + std::max(NTF, 1) * sizeof(int));
  const size_t o_V = off; off = align16(off + (size_t)d * d * 16);
  const size_t o_ph = off; off = align16(off + (size_t)std::max(T, 1) * d * 16);
  const size_t o_G = off; off = align16(off + (size_t)d * ld * 16);
  bool g_real = true;
  for (const cplx& gv : cp.G) g_real = g_real && gv.imag() == 0.0;
  const size_t n_gre = (size_t)(d / LC_T) * (d / LC_T) * LC_T * LC_GS;  // real table: tile-major with row stride LC_GS
  const size_t o_Gre = off; off = align16(off + (g_real ? n_gre * 8 : 16));
  const size_t o_pij = off; off = align16(off + pair_ij.size() * sizeof(int));
  const size_t o_ss = off; off = align16(off + cp.slot_start.size() * sizeof(int));
  const size_t o_av = off; off = align16(off + std::max<size_t>(cp.atom_val.size(), 1) * 16);
  const size_t o_ac = off; off = align16(off + std::max<size_t>(cp.atom_code.size(), 1) * sizeof(int));
  const size_t o_plan = off; off = align16(off + (size_t)npairs * LC_WARPS * 3 * sizeof(LcList));  // written by the plan kernel
  std::vector<unsigned char> blob(off, 0);
  if (NTF) {
    memcpy(&blob[o_tfsig], tf_sig.data(), NTF * sizeof(int));
    memcpy(&blob[o_tffreq], tf_freq.data(), NTF * sizeof(int));
  }
  memcpy(&blob[o_V], m->V.data(), (size_t)d * d * 16);
  {
    cplx* ph = reinterpret_cast<cplx*>(&blob[o_ph]);
    int slot = 0;
    for (size_t i = 0; i < g.out_times.size() && slot < T; i++)
      if (g.out_keep[i]) {
        for (int e = 0; e < d; e++) ph[(size_t)slot * d + e] = std::polar(1.0, -m->fe[e] * g.out_times[i]);
        slot++;
      }
  }
  memcpy(&blob[o_G], cp.G.data(), (size_t)d * ld * 16);
  if (g_real) {
    double* gr = reinterpret_cast<double*>(&blob[o_Gre]);
    const int nt = d / LC_T;
    for (int r = 0; r < d; r++)
      for (int c2 = 0; c2 < d; c2++) {
        const size_t tile = (size_t)(r / LC_T) * nt + c2 / LC_T;
        gr[tile * (LC_T * LC_GS) + (r % LC_T) * LC_GS + c2 % LC_T] = cp.G[tile * (LC_T * LC_TS) + (r % LC_T) * LC_TS + c2 % LC_T].real();
      }
  }
  memcpy(&blob[o_pij], pair_ij.data(), pair_ij.size() * sizeof(int));
  memcpy(&blob[o_ss], cp.slot_start.data(), cp.slot_start.size() * sizeof(int));
  if (!cp.atom_val.empty()) {
    memcpy(&blob[o_av], cp.atom_val.data(), cp.atom_val.size() * 16);
    memcpy(&blob[o_ac], cp.atom_code.data(), cp.atom_code.size() * sizeof(int));
  }
  const size_t n_per = (size_t)d * d, n_pad = (size_t)d * ld;
  if ((rc = m->tab_times.ensure(NT * sizeof(double)))) return rc;
  if ((rc = m->tab_idx.ensure(NT * sizeof(int)))) return rc;
  if ((rc = m->tab.ensure((size_t)NT * std::max(W, 2) * sizeof(double)))) return rc;
  if ((rc = m->work0.ensure(consts.size() * sizeof(double)))) return rc;
  if ((rc = m->work1.ensure(blob.size()))) return rc;
  // Gathers are unconditional (invalid columns carry a zero coefficient): the buffers are zero-filled once per solve, row
  // padding included, with slack in front of the first and behind the last matrix.
  const size_t slack = 2 * (size_t)LC_T * LC_TS;  // elements (two tiles): staged windows may run past the last tile
  const size_t w2_bytes = (4 * (size_t)B * n_pad + 2 * slack) * 16;
  if ((rc = m->work2.ensure(w2_bytes))) return rc;
  CASQ_CUDA(cudaMemsetAsync(m->work2.p, 0, w2_bytes, st));
  if ((rc = m->work3.ensure(3 * (size_t)B * cp.NC * 16))) return rc;
  CASQ_CUDA(cudaMemcpyAsync(m->tab_times.p, g.times.data(), NT * sizeof(double), cudaMemcpyHostToDevice, st));
  CASQ_CUDA(cudaMemcpyAsync(m->tab_idx.p, g.sample_idx.data(), NT * sizeof(int), cudaMemcpyHostToDevice, st));
  CASQ_CUDA(cudaMemcpyAsync(m->work0.p, consts.data(), consts.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  CASQ_CUDA(cudaMemcpyAsync(m->work1.p, blob.data(), blob.size(), cudaMemcpyHostToDevice, st));
  {
    const long long nthreads = NT * (KA + NF);
    sv_general_table_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>(
        (const double*)m->tab_times.p, NT, KA, NF, W, (const double*)m->work0.p, (const double*)m->work0.p + KA,
        (double*)m->tab.p);
    CASQ_CUDA(cudaGetLastError());
  }
  unsigned char* dev = (unsigned char*)m->work1.p;
  LbArgs a;
  memset(&a, 0, sizeof(a));
  a.d = d; a.ld = ld; a.KA = KA; a.NF = NF; a.NTF = NTF; a.W = W;
  a.tf_sig = (const int*)(dev + o_tfsig);
  a.tf_freq = (const int*)(dev + o_tffreq);
  a.table = (const double*)m->tab.p;
  a.idx = (const int*)m->tab_idx.p;
  a.params = params_dev;
  a.Bp = B;
  a.slots = ds;
  LcArgs c;
  memset(&c, 0, sizeof(c));
  c.d = d; c.ld = ld; c.nq = cp.nq; c.ntile = ntile; c.npairs = npairs; c.NC = cp.NC; c.NS = cp.NS;
  c.n2 = cp.n2; c.u_base = cp.u_base; c.w_base = cp.w_base;
  for (int k = 0; k < cp.n2; k++) {
    c.two_q[k] = cp.two_q[k];
    c.two_dir[k] = cp.two_dir[k];
  }
  c.drv_mask = cp.drv_mask;
  c.pair_ij = (const int*)(dev + o_pij);
  c.G = (const double2*)(dev + o_G);
  c.Gre = g_real ? (const double*)(dev + o_Gre) : nullptr;
  c.slot_start = (const int*)(dev + o_ss);
  c.atom_val = (const double2*)(dev + o_av);
  c.atom_code = (const int*)(dev + o_ac);
  c.plan = dev + o_plan;
  const size_t smem_coef = (size_t)std::max(NTF, 1) * 16;
  const bool low_only = (cp.drv_mask & 6u) == 0;  // only site 0 of the low three is driven (cfg 4): leaner in-tile code
  CASQ_CUDA(cudaFuncSetAttribute(lindblad_chain_coef_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_coef));
  // Two implementations of the same stage (lindblad_chain.cuh): operands through the LSU (LDG) or staged by bulk asynchronous
  // copies (TMA).  CASQ_LB_TMA_STAGES is a bit mask of the RK4 stages (bit 0 = stage 1) that use the TMA kernel.
  typedef void (*StageFn)(const LcArgs, int, const LbStage);
  StageFn fn_ldg = nullptr, fn_tma = nullptr;
#define LC_PICK(NQ_)                                                                                             \
  if (cp.nq == NQ_) {                                                                                            \
    fn_tma = low_only ? lindblad_chain_tma_kernel<NQ_, 1> : lindblad_chain_tma_kernel<NQ_, 7>;                   \
    fn_ldg = low_only ? lindblad_chain_stage_kernel<NQ_, 1> : lindblad_chain_stage_kernel<NQ_, 7>;               \
  }
  LC_PICK(3) LC_PICK(4) LC_PICK(5) LC_PICK(6)
#undef LC_PICK
  if (!fn_ldg) CASQ_FAIL(CASQ_ERR_UNSUPPORTED, "chain kernel: %d sites", cp.nq);
  const size_t smem_tma = lindblad_chain_tma_smem(cp.NC), smem_ldg = lindblad_chain_smem(cp.NC);
  CASQ_CUDA(cudaFuncSetAttribute(fn_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tma));
  CASQ_CUDA(cudaFuncSetAttribute(fn_ldg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ldg));
  {
    const dim3 pg((unsigned)npairs), pb(32);
    switch (cp.nq) {
      case 3: lindblad_chain_plan_kernel<3><<<pg, pb, 0, st>>>(c, (LcList*)(dev + o_plan)); break;
      case 4: lindblad_chain_plan_kernel<4><<<pg, pb, 0, st>>>(c, (LcList*)(dev + o_plan)); break;
      case 5: lindblad_chain_plan_kernel<5><<<pg, pb, 0, st>>>(c, (LcList*)(dev + o_plan)); break;
      default: lindblad_chain_plan_kernel<6><<<pg, pb, 0, st>>>(c, (LcList*)(dev + o_plan)); break;
    }
    CASQ_CUDA(cudaGetLastError());
  }
  if (getenv("CASQ_LB_CHECK_PLAN")) {
    // Self-check of the gather plan (compute-sanitizer is not available everywhere): every nine-row piece a block fetches,
    // from its own tiles or through a list offset, must lie inside one matrix plus the zero-filled slack around the
    // work buffers.  A piece past the end of a trajectory's matrix would read its neighbour's rows (harmless only while the
    // coefficient is zero), a piece outside the slack is an out-of-bounds access.
    std::vector<LcList> plan((size_t)npairs * LC_WARPS * 3);
    std::vector<int> pij(2 * (size_t)npairs);
    CASQ_CUDA(cudaStreamSynchronize(st));
    CASQ_CUDA(cudaMemcpy(plan.data(), dev + o_plan, plan.size() * sizeof(LcList), cudaMemcpyDeviceToHost));
    CASQ_CUDA(cudaMemcpy(pij.data(), dev + o_pij, pij.size() * sizeof(int), cudaMemcpyDeviceToHost));
    const long long TSZ = (long long)LC_T * LC_TS, piece = 9LL * LC_TS;
    long long lo = 0, hi = 0, worst_in = 0;
    for (int pr = 0; pr < npairs; pr++)
      for (int gq = 0; gq < LC_WARPS; gq++) {
        const long long I = pij[2 * pr], J = pij[2 * pr + 1];
        const long long tI = (I * ntile + J) * TSZ + 9LL * gq * LC_TS, tJ = (J * ntile + I) * TSZ + 9LL * gq * LC_TS;
        for (int seg = 0; seg < 3; seg++) {
          const LcList& L = plan[((size_t)pr * LC_WARPS + gq) * 3 + seg];
          if (L.n < 0 || L.n > 15) CASQ_FAIL(CASQ_ERR_INVALID, "chain plan: list (%d,%d,%d) has %d terms", pr, gq, seg, L.n);
          for (int i = 0; i < L.n; i++) {
            const long long start = (seg == 0 ? tJ : tI) + L.off[i], end = start + piece;
            lo = std::min(lo, start);
            hi = std::max(hi, end - (long long)n_pad);
            if (start >= 0 && end <= (long long)n_pad) worst_in++;
          }
        }
      }
    if (-lo > (long long)slack || hi > (long long)slack)
      CASQ_FAIL(CASQ_ERR_INVALID, "chain plan: a gather piece leaves the buffers (%lld elements before, %lld behind a matrix; slack %zu)", -lo,
                hi, slack);
    if (getenv("CASQ_LB_CHECK_PLAN_VERBOSE"))
      fprintf(stderr, "chain plan: %lld pieces inside their matrix, furthest %lld elements before / %lld behind (slack %zu)\n", worst_in, -lo, hi,
              slack);
  }
  int tma_stages = LC_DEFAULT_TMA_STAGES;
  if (const char* e = getenv("CASQ_LB_TMA_STAGES")) tma_stages = atoi(e) & 15;
  if (getenv("CASQ_LB_NO_TMA")) tma_stages = 0;

  struct Part {
    LbArgs a;
    LcArgs c;
    double2* bufs[4];
    long long b0, Bc;
    cudaStream_t st;
  };
  int n_parts = (int)std::min<long long>(4, std::max<long long>(1, B / 4));
  if (const char* e = getenv("CASQ_LB_PARTS")) n_parts = std::min(4, std::max(1, atoi(e)));
  n_parts = (int)std::min<long long>(n_parts, B);
  if (n_parts > 1 && !m->ev_fork) CASQ_CUDA(cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
  for (int i = 0; i + 1 < n_parts; i++)
    if (!m->aux_stream[i]) {
      CASQ_CUDA(cudaStreamCreateWithFlags(&m->aux_stream[i], cudaStreamNonBlocking));
      CASQ_CUDA(cudaEventCreateWithFlags(&m->ev_join[i], cudaEventDisableTiming));
    }
  Part parts[4];
  for (int pi = 0; pi < n_parts; pi++) {
    Part& P = parts[pi];
    P.b0 = (long long)B * pi / n_parts;
    P.Bc = (long long)B * (pi + 1) / n_parts - P.b0;
    P.st = pi == 0 ? st : m->aux_stream[pi - 1];
    P.a = a;
    P.a.B = P.Bc;
    P.a.b0 = P.b0;
    P.c = c;
    P.c.B = P.Bc;
    P.c.ctab = (double2*)m->work3.p + 3 * (size_t)P.b0 * cp.NC;
    for (int i = 0; i < 4; i++) P.bufs[i] = (double2*)m->work2.p + slack + (size_t)i * B * n_pad + (size_t)P.b0 * n_pad;
  }
  if (n_parts > 1) {
    CASQ_CUDA(cudaEventRecord(m->ev_fork, st));
    for (int i = 0; i + 1 < n_parts; i++) CASQ_CUDA(cudaStreamWaitEvent(m->aux_stream[i], m->ev_fork, 0));
  }
  for (int pi = 0; pi < n_parts; pi++) {
    Part& P = parts[pi];
    const long long n = P.Bc * (long long)n_per;
    const double2* r0p = (const double2*)rho0_dev + (rho0_batched ? (size_t)P.b0 * n_per : 0);
    lindblad_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, P.st>>>(d, ld, P.Bc, r0p, rho0_batched, P.bufs[0], tw, tpad);
    CASQ_CUDA(cudaGetLastError());
  }
  int t_slot = 0;
  auto emit = [&]() -> int {
    for (int pi = 0; pi < n_parts; pi++) {
      Part& P = parts[pi];
      if (out_rho_dev) {
        const long long n = P.Bc * (long long)n_per;
        lindblad_copy_out_kernel<<<(unsigned)((n + 255) / 256), 256, 0, P.st>>>(d, ld, P.Bc, P.bufs[0],
                                                                              (double2*)out_rho_dev + (size_t)P.b0 * T * n_per, T, t_slot, tw, tpad);
        CASQ_CUDA(cudaGetLastError());
      }
      if (out_pops_dev) {
        lindblad_pops_kernel<<<dim3((unsigned)d, (unsigned)P.Bc), 256, 0, P.st>>>(d, ld, P.bufs[0], (const double2*)(dev + o_V),
                                                                                (const double2*)(dev + o_ph) + (size_t)t_slot * d,
                                                                                out_pops_dev + (size_t)P.b0 * T * d, T, t_slot, tw, tpad);
        CASQ_CUDA(cudaGetLastError());
      }
    }
    t_slot++;
    return CASQ_OK;
  };
  if (g.out_keep[0] && (rc = emit())) return rc;
  long long pos = 0;
  for (size_t p = 0; p < g.piece_nsteps.size(); p++) {
    const int n = g.piece_nsteps[p];
    const double h_piece = g.piece_h[p];
    for (int s = 0; s < n; s++) {
      for (int pi = 0; pi < n_parts; pi++) {
        Part& P = parts[pi];
        const long long tasks = P.Bc * (long long)npairs;
        const unsigned grid = (unsigned)tasks;  // one block of three warps per tile pair
        lindblad_chain_coef_kernel<<<dim3((unsigned)P.Bc, 3), 128, smem_coef, P.st>>>(P.a, P.c, pos);
        for (int stage = 1; stage <= 4; stage++) {
          const int set = stage == 1 ? 0 : (stage == 4 ? 2 : 1);
          LbStage sg;
          sg.in = P.bufs[stage - 1];
          sg.y = P.bufs[0];
          sg.y1 = P.bufs[1];
          sg.y2 = P.bufs[2];
          sg.out = stage == 4 ? P.bufs[0] : P.bufs[stage];
          sg.stage = stage;
          sg.h = h_piece;
          if ((tma_stages >> (stage - 1)) & 1)
            fn_tma<<<grid, 32 * LC_WARPS, smem_tma, P.st>>>(P.c, set, sg);
          else
            fn_ldg<<<grid, 32 * LC_WARPS, smem_ldg, P.st>>>(P.c, set, sg);
        }
      }
      pos += 2;
    }
    CASQ_CUDA(cudaGetLastError());
    pos += 1;
    if (g.out_keep[p + 1] && (rc = emit())) return rc;
  }
  for (int i = 0; i + 1 < n_parts; i++) {
    CASQ_CUDA(cudaEventRecord(m->ev_join[i], m->aux_stream[i]));
    CASQ_CUDA(cudaStreamWaitEvent(st, m->ev_join[i], 0));
  }
  return CASQ_OK;
}
}  // namespace

extern "C" int casq_solve_lindblad_rk4(casq_model* m, int n_slots, const casq_pulse_slot* slots, int64_t B,
                                       const double* params_dev, const double* rho0_dev, int rho0_batched, double t0, double t1,
                                       double max_dt, int T, const double* t_eval, double* out_rho_dev, double* out_pops_dev,
                                       void* stream) {
  if (!m) CASQ_FAIL(CASQ_ERR_INVALID, "casq_solve_lindblad_rk4: NULL model");
  if (B < 0 || !rho0_dev || (n_slots > 0 && !params_dev)) CASQ_FAIL(CASQ_ERR_INVALID, "casq_solve_lindblad_rk4: NULL buffer");
  if (!out_rho_dev && !out_pops_dev) CASQ_FAIL(CASQ_ERR_INVALID, "casq_solve_lindblad_rk4: nothing to output");
  if (!(t1 > t0)) CASQ_FAIL(CASQ_ERR_INVALID, "casq_solve_lindblad_rk4: need t1 > t0");
  if (!m->U_is_identity)
    CASQ_FAIL(CASQ_ERR_UNSUPPORTED, "casq_solve_lindblad_rk4 needs a diagonal rotating frame (1-D frame or none): the structured "
                                    "superoperator works in the bare tensor-product basis");
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
  if (KA > LB_MAXK) CASQ_FAIL(CASQ_ERR_UNSUPPORTED, "at most %d driven channels", LB_MAXK);
  const int d = m->d;
  if (d > 1024) CASQ_FAIL(CASQ_ERR_UNSUPPORTED, "dim %d too large", d);
  cudaStream_t st = (cudaStream_t)stream;
  CASQ_CUDA(cudaSetDevice(m->device));

  // ---------------- host: shift-class tables
  Builder bld;
  bld.d = d;
  double scale = 0;
  for (int i = 0; i < d; i++) scale = std::max(scale, std::fabs(m->fe[i]));
  bld.tol = 2e-14 * (scale > 0 ? scale : 1.0);  // a few ulps of the largest frame eigenvalue: merges e_i - e_j that differ by rounding only
  const cplx mi(0.0, -1.0);
  std::vector<Atom> one;
  auto bohr = [&](int i, int j) { return m->fe[i] - m->fe[j]; };
  for (int i = 0; i < d; i++)
    for (int j = 0; j < d; j++) {
      if (std::abs(m->S[i * d + j]) != 0.0) one.push_back({i, j, 0, mi * m->S[i * d + j], bohr(i, j)});
      for (int k = 0; k < KA; k++) {
        const cplx p = m->P[((size_t)active[k] * d + i) * d + j], n = m->N[((size_t)active[k] * d + i) * d + j];
        if (!m->rwa) {
          if (std::abs(p) != 0.0) one.push_back({i, j, 1 + 2 * KA + k, mi * (2.0 * p), bohr(i, j)});  // Re(z_k) * G
        } else {
          if (std::abs(p) != 0.0) one.push_back({i, j, 1 + k, mi * p, bohr(i, j)});
          if (std::abs(n) != 0.0) one.push_back({i, j, 1 + KA + k, mi * n, bohr(i, j)});
        }
      }
    }
  // D = sum_j L_j^+ L_j
  std::vector<cplx> D((size_t)d * d, 0.0);
  for (int j = 0; j < m->J; j++) {
    const cplx* L = &m->Lfb[(size_t)j * d * d];
    for (int a_ = 0; a_ < d; a_++)
      for (int k = 0; k < d; k++) {
        const cplx lka = std::conj(L[k * d + a_]);
        if (std::abs(lka) == 0.0) continue;
        for (int b_ = 0; b_ < d; b_++) D[a_ * d + b_] += lka * L[k * d + b_];
      }
  }
  for (int i = 0; i < d; i++)
    for (int j = 0; j < d; j++)
      if (std::abs(D[i * d + j]) != 0.0) one.push_back({i, j, 0, -0.5 * D[i * d + j], bohr(i, j)});
  if (B > 65535) CASQ_FAIL(CASQ_ERR_UNSUPPORTED, "casq_solve_lindblad_rk4: at most 65535 density matrices per call");
  {
    // chains of three-level sites (every casq TransmonModel with >= 3 qubits) take the register-tiled kernel
    Builder probe = bld;
    const ChainPlan cp = lc_analyze(m, probe, one);
    if (cp.ok)
      return lindblad_run_chain(m, cp, probe, KA, active, ds, n_total, B, params_dev, rho0_dev, rho0_batched, t0, t1, max_dt, T,
                                t_eval, out_rho_dev, out_pops_dev, st);
  }
  ClassTab os;
  bld.place(one, &os);
  if (os.delta.empty()) {  // completely free evolution: keep one empty class so the kernel has something to index
    os.delta.push_back(0);
    os.code.assign(d, (short)-1);
    os.val.assign(d, cplx(0.0));
  }
  ClassTab us;
  std::vector<int> pairs;
  for (int j = 0; j < m->J; j++) {
    const cplx* L = &m->Lfb[(size_t)j * d * d];
    std::vector<Atom> atoms;
    for (int a_ = 0; a_ < d; a_++)
      for (int b_ = 0; b_ < d; b_++)
        if (std::abs(L[a_ * d + b_]) != 0.0) atoms.push_back({a_, b_, 0, L[a_ * d + b_], bohr(a_, b_)});
    const int first = (int)us.delta.size();
    ClassTab mine;
    bld.place(atoms, &mine);
    for (size_t c = 0; c < mine.delta.size(); c++) {
      us.delta.push_back(mine.delta[c]);
      us.code.insert(us.code.end(), mine.code.begin() + c * d, mine.code.begin() + (c + 1) * d);
      us.val.insert(us.val.end(), mine.val.begin() + c * d, mine.val.begin() + (c + 1) * d);
    }
    for (size_t c1 = 0; c1 < mine.delta.size(); c1++)
      for (size_t c2 = 0; c2 < mine.delta.size(); c2++) {
        pairs.push_back(first + (int)c1);
        pairs.push_back(first + (int)c2);
      }
  }
  const int NQ = (int)os.delta.size(), NU = (int)us.delta.size(), NP = (int)pairs.size() / 2;
  const int NF = (int)bld.freqs.size(), NTF = (int)bld.tfs.size();
  if (NTF > 30000) CASQ_FAIL(CASQ_ERR_UNSUPPORTED, "too many distinct time functions (%d)", NTF);

  // group the two-sided class pairs by their shifts: members of a group gather the same rho element
  std::vector<int> grp_start, grp_pairs;
  std::vector<cplx> cdense;  // [d*d] folded constant diagonal two-sided coefficients (already x 1/2), or empty
  {
    std::map<std::pair<int, int>, std::vector<int>> groups;
    for (int p = 0; p < NP; p++) groups[{us.delta[pairs[2 * p]], us.delta[pairs[2 * p + 1]]}].push_back(p);
    grp_start.push_back(0);
    for (auto& kv : groups) {
      // A group that gathers the element itself (delta1 = delta2 = 0) with time-independent factors (e.g. the five
      // pure-dephasing operators n_q) is folded into ONE dense constant coefficient matrix: 1 load instead of 1 per pair.
      if (kv.first.first == 0 && kv.first.second == 0) {
        bool constant = true;
        for (int p : kv.second)
          for (int side = 0; side < 2 && constant; side++) {
            const int q = pairs[2 * p + side];
            for (int r = 0; r < d && constant; r++) {
              const int code = us.code[(size_t)q * d + r];
              if (code >= 0 && (bld.tfs[code].first != 0 || bld.tfs[code].second != -1)) constant = false;
            }
          }
        if (constant) {
          if (cdense.empty()) cdense.assign((size_t)d * d, cplx(0.0));
          for (int p : kv.second) {
            const int q1 = pairs[2 * p], q2 = pairs[2 * p + 1];
            for (int r = 0; r < d; r++) {
              if (us.code[(size_t)q1 * d + r] < 0) continue;
              const cplx ul = 0.5 * us.val[(size_t)q1 * d + r];
              for (int c = 0; c < d; c++)
                if (us.code[(size_t)q2 * d + c] >= 0) cdense[(size_t)r * d + c] += ul * std::conj(us.val[(size_t)q2 * d + c]);
            }
          }
          continue;
        }
      }
      for (int p : kv.second) {
        grp_pairs.push_back(pairs[2 * p]);
        grp_pairs.push_back(pairs[2 * p + 1]);
      }
      grp_start.push_back((int)grp_pairs.size() / 2);
    }
  }
  const int NG = (int)grp_start.size() - 1;

  // ---------------- time grid and stage table (carriers + distinct Bohr phases)
  TimeGrid g;
  rc = casq_build_time_grid(t0, t1, max_dt, T, t_eval, m->dt, n_total, &g);
  if (rc != CASQ_OK) return rc;
  const long long NT = (long long)g.times.size();
  const int W = 2 * KA + 2 * NF;
  m->tab_key.clear();
  std::vector<double> consts(KA + std::max(NF, 1));
  for (int k = 0; k < KA; k++) consts[k] = 2.0 * M_PI * m->nu[active[k]];
  for (int f = 0; f < NF; f++) consts[KA + f] = bld.freqs[f];

  // device tables, packed into work1
  std::vector<int> tf_sig(NTF), tf_freq(NTF);
  for (int n = 0; n < NTF; n++) {
    tf_sig[n] = bld.tfs[n].first;
    tf_freq[n] = bld.tfs[n].second;
  }
  auto align16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
  size_t off = 0;
  const size_t o_qdelta = off; off = align16(off + NQ * sizeof(int));
  const size_t o_qcode = off; off = align16(off + (size_t)NQ * d * sizeof(short));
  const size_t o_qval = off; off = align16(off + (size_t)NQ * d * 16);
  const size_t o_udelta = off; off = align16(off + std::max(NU, 1) * sizeof(int));
  const size_t o_ucode = off; off = align16(off + (size_t)std::max(NU, 1) * d * sizeof(short));
  const size_t o_uval = off; off = align16(off + (size_t)std::max(NU, 1) * d * 16);
  const size_t o_pairs = off; off = align16(off + std::max(NP, 1) * 2 * sizeof(int));
  const size_t o_gstart = off; off = align16(off + (NG + 1) * sizeof(int));
  const size_t o_tfsig = off; off = align16(off + std::max(NTF, 1) * sizeof(int));
  const size_t o_tffreq = off; off = align16(off + std::max(NTF, 1) * sizeof(int));
  const size_t o_V = off; off = align16(off + (size_t)d * d * 16);
  const size_t o_ph = off; off = align16(off + (size_t)std::max(T, 1) * d * 16);  // e^{-i e t} for every output time
  const int ld = (d + 7) & ~7;  // 128 B aligned rows: a warp's 512 B gather touches 4 lines, not 5
  const size_t o_cd = off; off = align16(off + (cdense.empty() ? 16 : (size_t)d * ld * 16));
  const int ntile = (d + LB_TILE - 1) / LB_TILE;
  std::vector<int> pair_ij;
  for (int I = 0; I < ntile; I++)  // off-diagonal pairs (twice the work) first
    for (int J = I + 1; J < ntile; J++) {
      pair_ij.push_back(I);
      pair_ij.push_back(J);
    }
  for (int I = 0; I < ntile; I++) {
    pair_ij.push_back(I);
    pair_ij.push_back(I);
  }
  const size_t o_pij = off; off = align16(off + pair_ij.size() * sizeof(int));
  std::vector<unsigned char> blob(off, 0);
  memcpy(&blob[o_qdelta], os.delta.data(), NQ * sizeof(int));
  memcpy(&blob[o_qcode], os.code.data(), (size_t)NQ * d * sizeof(short));
  memcpy(&blob[o_qval], os.val.data(), (size_t)NQ * d * 16);
  if (NU) {
    memcpy(&blob[o_udelta], us.delta.data(), NU * sizeof(int));
    memcpy(&blob[o_ucode], us.code.data(), (size_t)NU * d * sizeof(short));
    memcpy(&blob[o_uval], us.val.data(), (size_t)NU * d * 16);
  }
  if (!grp_pairs.empty()) memcpy(&blob[o_pairs], grp_pairs.data(), grp_pairs.size() * sizeof(int));
  memcpy(&blob[o_gstart], grp_start.data(), (NG + 1) * sizeof(int));
  if (NTF) {
    memcpy(&blob[o_tfsig], tf_sig.data(), NTF * sizeof(int));
    memcpy(&blob[o_tffreq], tf_freq.data(), NTF * sizeof(int));
  }
  memcpy(&blob[o_V], m->V.data(), (size_t)d * d * 16);
  {
    cplx* ph = reinterpret_cast<cplx*>(&blob[o_ph]);
    int slot = 0;
    for (size_t i = 0; i < g.out_times.size() && slot < T; i++)
      if (g.out_keep[i]) {
        for (int e = 0; e < d; e++) ph[(size_t)slot * d + e] = std::polar(1.0, -m->fe[e] * g.out_times[i]);
        slot++;
      }
  }
  if (!cdense.empty())
    for (int r = 0; r < d; r++) memcpy(&blob[o_cd + (size_t)r * ld * 16], &cdense[(size_t)r * d], (size_t)d * 16);
  memcpy(&blob[o_pij], pair_ij.data(), pair_ij.size() * sizeof(int));

  const size_t n_per = (size_t)d * d;
  if ((rc = m->tab_times.ensure(NT * sizeof(double)))) return rc;
  if ((rc = m->tab_idx.ensure(NT * sizeof(int)))) return rc;
  if ((rc = m->tab.ensure((size_t)NT * std::max(W, 2) * sizeof(double)))) return rc;
  if ((rc = m->work0.ensure(consts.size() * sizeof(double)))) return rc;
  if ((rc = m->work1.ensure(blob.size()))) return rc;
  if (NQ > LB_MAXQ) CASQ_FAIL(CASQ_ERR_UNSUPPORTED, "more than %d one-sided shift classes (%d): operators too dense for the structured path", LB_MAXQ, NQ);
  const size_t n_pad = (size_t)d * ld;
  if ((rc = m->work2.ensure(4 * (size_t)B * n_pad * 16))) return rc;
  const size_t n_cq = (size_t)B * NQ * d, n_cu = (size_t)B * std::max(NU, 1) * d;
  if ((rc = m->work3.ensure(3 * (n_cq + n_cu) * 16))) return rc;
  double2* cq_all = (double2*)m->work3.p;
  double2* cu_all = cq_all + 3 * n_cq;
  // compacted per-row lists (3 stage times x B x d rows)
  const size_t n_rows = 3 * (size_t)B * d;
  const size_t lists_bytes = n_rows * ((size_t)NQ * (16 + 4) + (size_t)std::max(NG, 1) * 4 + 8) + 64;
  if ((rc = m->pieces_h.ensure(lists_bytes))) return rc;
  LbLists lists;
  {
    unsigned char* lp = (unsigned char*)m->pieces_h.p;
    lists.lcoef = (double2*)lp; lp += n_rows * NQ * 16;
    lists.loff = (int*)lp; lp += n_rows * NQ * 4;
    lists.glist = (int*)lp; lp += n_rows * std::max(NG, 1) * 4;
    lists.lcnt = (int*)lp; lp += n_rows * 4;
    lists.gcnt = (int*)lp;
  }
  CASQ_CUDA(cudaMemcpyAsync(m->tab_times.p, g.times.data(), NT * sizeof(double), cudaMemcpyHostToDevice, st));
  CASQ_CUDA(cudaMemcpyAsync(m->tab_idx.p, g.sample_idx.data(), NT * sizeof(int), cudaMemcpyHostToDevice, st));
  CASQ_CUDA(cudaMemcpyAsync(m->work0.p, consts.data(), consts.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  CASQ_CUDA(cudaMemcpyAsync(m->work1.p, blob.data(), blob.size(), cudaMemcpyHostToDevice, st));
  {
    const long long nthreads = NT * (KA + NF);
    sv_general_table_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>(
        (const double*)m->tab_times.p, NT, KA, NF, W, (const double*)m->work0.p, (const double*)m->work0.p + KA,
        (double*)m->tab.p);
    CASQ_CUDA(cudaGetLastError());
  }
  unsigned char* dev = (unsigned char*)m->work1.p;
  if (B > 65535) CASQ_FAIL(CASQ_ERR_UNSUPPORTED, "casq_solve_lindblad_rk4: at most 65535 density matrices per call");
  LbArgs a;
  memset(&a, 0, sizeof(a));
  a.d = d; a.ld = ld;
  a.KA = KA; a.NF = NF; a.NTF = NTF; a.NQ = NQ; a.NU = NU; a.NG = NG; a.NP = (int)grp_pairs.size() / 2; a.W = W;
  a.grp_start = (const int*)(dev + o_gstart);
  a.q_delta = (const int*)(dev + o_qdelta);
  a.q_code = (const short*)(dev + o_qcode);
  a.q_val = (const double2*)(dev + o_qval);
  a.u_delta = (const int*)(dev + o_udelta);
  a.u_code = (const short*)(dev + o_ucode);
  a.u_val = (const double2*)(dev + o_uval);
  a.grp_pairs = (const int*)(dev + o_pairs);
  a.tf_sig = (const int*)(dev + o_tfsig);
  a.tf_freq = (const int*)(dev + o_tffreq);
  a.table = (const double*)m->tab.p;
  a.idx = (const int*)m->tab_idx.p;
  a.params = params_dev;
  a.Bp = B;
  a.slots = ds;
  a.cdense = cdense.empty() ? nullptr : (const double2*)(dev + o_cd);
  a.pair_ij = (const int*)(dev + o_pij);
  if (NU > LB_MAXQ || NG > LB_MAXQ || 2 * NP > 4 * LB_MAXQ)
    CASQ_FAIL(CASQ_ERR_UNSUPPORTED, "too many two-sided classes for the structured Lindblad kernel (%d classes, %d pairs)", NU, NP);
  const size_t smem_coef = (size_t)std::max(NTF, 1) * 16;
  CASQ_CUDA(cudaFuncSetAttribute(lindblad_coef_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_coef));
  const size_t smem_stage = lindblad_stage_smem(NQ, NU, NG);
  CASQ_CUDA(cudaFuncSetAttribute(lindblad_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_stage));

  // The batch runs as two (CASQ_LB_PARTS: up to four) parts on as many streams: every stage is its own grid and ends in a partly filled wave
  // (4.3 waves at B = 128), and the blocks of the other parts fill that tail.  Each part has its own slice of every
  // per-trajectory workspace; the tables are shared.
  struct Part {
    LbArgs a;
    LbLists lists;
    double2 *cq, *cu, *bufs[4];
    long long b0, Bc;
    cudaStream_t st;
  };
  // measured at d = 243: B = 128: 161 / 178 / 183 / 184 traj/s for 1 / 2 / 3 / 4 parts; B = 16: 84 / 95 / - / 97.  Small
  // matrices (a stage is a few microseconds) stay on one stream: four times the launches would only add host time.
  int n_parts = d >= 64 ? (int)std::min<long long>(4, std::max<long long>(1, B / 4)) : 1;
  if (const char* e = getenv("CASQ_LB_PARTS")) n_parts = std::min(4, std::max(1, atoi(e)));  // tuning knob
  n_parts = (int)std::min<long long>(n_parts, B);
  if (n_parts > 1 && !m->ev_fork) CASQ_CUDA(cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
  for (int i = 0; i + 1 < n_parts; i++)
    if (!m->aux_stream[i]) {
      CASQ_CUDA(cudaStreamCreateWithFlags(&m->aux_stream[i], cudaStreamNonBlocking));
      CASQ_CUDA(cudaEventCreateWithFlags(&m->ev_join[i], cudaEventDisableTiming));
    }
  Part parts[4];
  const size_t nu1 = (size_t)std::max(NU, 1), ng1 = (size_t)std::max(NG, 1);
  for (int pi = 0; pi < n_parts; pi++) {
    Part& P = parts[pi];
    P.b0 = (long long)B * pi / n_parts;
    P.Bc = (long long)B * (pi + 1) / n_parts - P.b0;
    P.st = pi == 0 ? st : m->aux_stream[pi - 1];
    P.a = a;
    P.a.B = P.Bc;
    P.a.b0 = P.b0;
    for (int i = 0; i < 4; i++) P.bufs[i] = (double2*)m->work2.p + (size_t)i * B * n_pad + (size_t)P.b0 * n_pad;
    P.cq = cq_all + 3 * (size_t)P.b0 * NQ * d;
    P.cu = cu_all + 3 * (size_t)P.b0 * nu1 * d;
    const size_t r0 = 3 * (size_t)P.b0 * d;  // first row of this part in every list array
    P.lists.lcoef = lists.lcoef + r0 * NQ;
    P.lists.loff = lists.loff + r0 * NQ;
    P.lists.glist = lists.glist + r0 * ng1;
    P.lists.lcnt = lists.lcnt + r0;
    P.lists.gcnt = lists.gcnt + r0;
  }
  if (n_parts > 1) {  // fork: the other streams start after the table uploads and the table kernel
    CASQ_CUDA(cudaEventRecord(m->ev_fork, st));
    for (int i = 0; i + 1 < n_parts; i++) CASQ_CUDA(cudaStreamWaitEvent(m->aux_stream[i], m->ev_fork, 0));
  }
  for (int pi = 0; pi < n_parts; pi++) {
    Part& P = parts[pi];
    const long long n = P.Bc * (long long)n_per;
    const double2* r0p = (const double2*)rho0_dev + (rho0_batched ? (size_t)P.b0 * n_per : 0);
    lindblad_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, P.st>>>(d, ld, P.Bc, r0p, rho0_batched, P.bufs[0]);
    CASQ_CUDA(cudaGetLastError());
  }
  int t_slot = 0;
  auto emit = [&]() -> int {
    for (int pi = 0; pi < n_parts; pi++) {
      Part& P = parts[pi];
      if (out_rho_dev) {
        const long long n = P.Bc * (long long)n_per;
        lindblad_copy_out_kernel<<<(unsigned)((n + 255) / 256), 256, 0, P.st>>>(d, ld, P.Bc, P.bufs[0],
                                                                              (double2*)out_rho_dev + (size_t)P.b0 * T * n_per, T, t_slot);
        CASQ_CUDA(cudaGetLastError());
      }
      if (out_pops_dev) {
        lindblad_pops_kernel<<<dim3((unsigned)d, (unsigned)P.Bc), 256, 0, P.st>>>(d, ld, P.bufs[0], (const double2*)(dev + o_V),
                                                                                (const double2*)(dev + o_ph) + (size_t)t_slot * d,
                                                                                out_pops_dev + (size_t)P.b0 * T * d, T, t_slot);
        CASQ_CUDA(cudaGetLastError());
      }
    }
    t_slot++;
    return CASQ_OK;
  };
  if (g.out_keep[0] && (rc = emit())) return rc;
  long long pos = 0;
  for (size_t p = 0; p < g.piece_nsteps.size(); p++) {
    const int n = g.piece_nsteps[p];
    const double h_piece = g.piece_h[p];
    for (int s = 0; s < n; s++) {
      // launches of the two parts are interleaved step by step so that neither stream's queue runs dry
      for (int pi = 0; pi < n_parts; pi++) {
        Part& P = parts[pi];
        const dim3 grid_pairs((unsigned)(pair_ij.size() / 2), (unsigned)P.Bc);
        lindblad_coef_kernel<<<dim3((unsigned)P.Bc, 3), 256, smem_coef, P.st>>>(P.a, pos, P.cq, P.cu, P.lists);
        // stage 1: in = y (bufs[0]) -> y1 (bufs[1]); 2: in = y1 -> y2; 3: in = y2 -> y3; 4: in = y3 -> y (in place)
        for (int stage = 1; stage <= 4; stage++) {
          const int set = stage == 1 ? 0 : (stage == 4 ? 2 : 1);
          LbStage sg;
          sg.in = P.bufs[stage - 1];
          sg.y = P.bufs[0];
          sg.y1 = P.bufs[1];
          sg.y2 = P.bufs[2];
          sg.out = stage == 4 ? P.bufs[0] : P.bufs[stage];
          sg.stage = stage;
          sg.h = h_piece;
          lindblad_stage_kernel<<<grid_pairs, 512, smem_stage, P.st>>>(P.a, P.lists, P.cu, set, sg);
        }
      }
      pos += 2;
    }
    CASQ_CUDA(cudaGetLastError());
    pos += 1;
    if (g.out_keep[p + 1] && (rc = emit())) return rc;
  }
  for (int i = 0; i + 1 < n_parts; i++) {  // join
    CASQ_CUDA(cudaEventRecord(m->ev_join[i], m->aux_stream[i]));
    CASQ_CUDA(cudaStreamWaitEvent(st, m->ev_join[i], 0));
  }
  return CASQ_OK;
}
