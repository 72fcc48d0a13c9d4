// casq_b200: Lindblad RK4 stage kernel for CHAINS OF THREE-LEVEL SITES (BASELINE cfg 4: the five-transmon ibmq_manila model,
// d = 3^5 = 243).  casq's TransmonModel always builds this structure: three levels per qubit (src/casq/models/
// transmon_model.py:150), single-site ladder drives X_q, nearest-neighbour exchange couplings, and the T1/T2 noise model
// adds single-site lowering operators and diagonal dephasing operators.  lindblad.cu detects the structure from the dense
// operators handed over the C ABI (every element must fit; anything else runs the generic shift-class kernel of
// lindblad_stage.cuh) and this kernel exploits it:
//
//   index r = sum_q x_q 3^q.   One WARP owns the 27 x 27 tile pair (I,J), (J,I), I <= J, of one trajectory; lane = column
//   inside the tile, and a lane keeps the 27 accumulators of its column (the digits x_0, x_1, x_2 of the row index) in
//   REGISTERS.  Every operator that acts inside the low three sites (X_0..X_2, J_01, J_12: 4.1 of the 5.9 non-zeros per row
//   of A = -iH_F - D_F/2 for ibmq_manila) is then applied from ONE load of each source element: a loaded element feeds
//   ~4 complex FMAs instead of one (the generic kernel gathers once per non-zero: ~100 L1 wavefronts per 32 outputs, the
//   bound ncu showed in profiles/r01_ncu_lindblad_fused_stage.txt).  Operators reaching outside the tile (J_23, J_34, X_3..,
//   and the two-sided terms L rho L^+) are "gather terms": one extra row load each, with warp-uniform or per-lane
//   coefficients; all address offsets are immediates.
//   k = M + M^+ + sum_j L_j rho L_j^+ with M = A rho is formed as in the generic kernel: M_JI is computed with left gathers,
//   transposed through a per-warp shared tile, the RK4 combination is fused, and the mirror tile out[J,I] = out[I,J]^+ is
//   written from the shared tile, so every stage output is exactly Hermitian and y[J,I], y1[J,I].. are never read.
//   Coefficients (value x time function) are a 1 kB table per (trajectory, stage time) written once per step by
//   lindblad_chain_coef_kernel -- the generic path's per-row lists (tens of MB per step) are gone.
#pragma once
#include "lindblad_stage.cuh"

#define LC_T 27          // tile edge = 3 sites
#define LC_WARPS 8       // warps (tile pairs) per block
#define LC_MAXQ 6        // sites
#define LC_MAX2 8        // two-sided single-site terms
#define LC_MAXTERMS 32   // gather terms per task

// coefficient-table layout per (stage set, trajectory): X slots, J slots, U slots, W slots
//   X(q, dir, n)      = ((q*2 + dir)*2 + n)                    dir 0: A[r, r+3^q], n = x_q(r);  dir 1: A[r, r-3^q], n = x_q(r)-1
//   J(q, dir, n, m)   = 4nq + (((q*2 + dir)*2 + n)*2 + m)      dir 0: A[r, r+3^(q+1)-3^q], n = x_q(r)-1, m = x_(q+1)(r)
//                                                             dir 1: A[r, r-3^(q+1)+3^q], n = x_q(r),   m = x_(q+1)(r)-1
//   U(term, n)        = u_base + term*2 + n                    L[r, r+dir*3^q], n = x_q(r) (dir +1) or x_q(r)-1 (dir -1)
//   W(term, nr, nc)   = w_base + (term*2 + nr)*2 + nc          U(term,nr) conj(U(term,nc))
struct LcArgs {
  int d, ld, nq, ntile, npairs, NC, NS;
  long long B;  // trajectories of this part
  int n2, u_base, w_base;
  int two_q[LC_MAX2], two_dir[LC_MAX2];
  unsigned drv_mask;  // sites with a non-empty X slot
  const int* pair_ij;         // [npairs][2]
  double2* ctab;              // [3][B][NC]
  const double2* G;           // [d][ld]: A_rr + conj(A_cc) + sum over diagonal L_j of l_j(r) conj(l_j(c))
  // coefficient kernel inputs: slot s = sum over atoms [slot_start[s], slot_start[s+1]) of atom_val * tf[atom_code]
  const int* slot_start;
  const double2* atom_val;
  const int* atom_code;
};

__device__ __forceinline__ void lc_cfma(double2& acc, const double2 c, const double2 v) {
  acc.x = fma(c.x, v.x, fma(-c.y, v.y, acc.x));
  acc.y = fma(c.x, v.y, fma(c.y, v.x, acc.y));
}

// per (trajectory, stage time): time functions -> slot coefficients
__global__ void __launch_bounds__(128) lindblad_chain_coef_kernel(const LbArgs a, const LcArgs c, long long entry0) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* s_tf = reinterpret_cast<double2*>(smem_raw);  // [NTF]
  __shared__ dcplx s_env[LB_MAXK];
  const long long b = blockIdx.x;
  const int set = blockIdx.y, tid = threadIdx.x;
  const long long entry = entry0 + set;
  if (tid == 0) {
    dcplx env[LB_MAXK];
    casq_channel_envelopes<LB_MAXK>(a.slots, a.params, a.Bp, a.b0 + b, __ldg(a.idx + entry), env);
#pragma unroll
    for (int k = 0; k < LB_MAXK; k++) s_env[k] = env[k];
  }
  __syncthreads();
  const double* ent = a.table + entry * a.W;
  for (int n = tid; n < a.NTF; n += blockDim.x) {
    const int sig = a.tf_sig[n], f = a.tf_freq[n];
    double2 v = make_double2(1.0, 0.0);
    if (sig > 0) {
      const int k = (sig - 1) % a.KA, kind = (sig - 1) / a.KA;
      const double cc = ent[2 * k], ss = ent[2 * k + 1];
      const double zr = s_env[k].re * cc - s_env[k].im * ss, zi = s_env[k].re * ss + s_env[k].im * cc;
      v = kind == 0 ? make_double2(zr, zi) : (kind == 1 ? make_double2(zr, -zi) : make_double2(zr, 0.0));
    }
    if (f >= 0) v = lb_cmul(v, make_double2(ent[2 * a.KA + 2 * f], ent[2 * a.KA + 2 * f + 1]));
    s_tf[n] = v;
  }
  __syncthreads();
  double2* out = c.ctab + ((long long)set * c.B + b) * c.NC;
  for (int s = tid; s < c.NS; s += blockDim.x) {
    double2 acc = make_double2(0.0, 0.0);
    for (int i = c.slot_start[s]; i < c.slot_start[s + 1]; i++) lc_cfma(acc, c.atom_val[i], s_tf[c.atom_code[i]]);
    out[s] = acc;
  }
  __syncthreads();  // U slots written by this block
  for (int i = tid; i < 4 * c.n2; i += blockDim.x) {
    const int term = i >> 2, nr = (i >> 1) & 1, nc = i & 1;
    const double2 u = out[c.u_base + term * 2 + nr], w = out[c.u_base + term * 2 + nc];
    out[c.w_base + i] = make_double2(u.x * w.x + u.y * w.y, u.y * w.x - u.x * w.y);  // u conj(w)
  }
}

__host__ __device__ constexpr int lc_pow3(int q) { return q == 0 ? 1 : 3 * lc_pow3(q - 1); }

// The in-tile part of M = A rho for the 27 rows of a row tile at this lane's column: every source row is loaded once and
// applied to all the accumulators it feeds (X_0..X_2 ladder elements, J_01 and J_12 exchange elements).  ct = the warp's
// coefficient table in shared memory (broadcast reads).  XLOW: bit q set = the ladder slots of site q are compiled in.
template <int NQ, int LD, int XLOW>
__device__ __forceinline__ void lc_in_tile(double2 (&acc)[LC_T], const double2* __restrict__ p, const double2* ct) {
  const double2* cJ = ct + 4 * NQ;
#pragma unroll
  for (int s = 0; s < LC_T; s++) {  // source row s of the tile
    const double2 v = __ldg(p + s * LD);
    const int xs[3] = {s % 3, (s / 3) % 3, s / 9};
    const int pw[3] = {1, 3, 9};
#pragma unroll
    for (int q = 0; q < 3; q++) {
      if (!((XLOW >> q) & 1)) continue;
      // A[t, s] with s = t + 3^q: slot X(q, 0, n = x_q(t) = x_q(s) - 1)
      if (xs[q] >= 1) lc_cfma(acc[s - pw[q]], ct[(q * 2 + 0) * 2 + xs[q] - 1], v);
      // A[t, s] with s = t - 3^q: slot X(q, 1, n = x_q(t) - 1 = x_q(s))
      if (xs[q] <= 1) lc_cfma(acc[s + pw[q]], ct[(q * 2 + 1) * 2 + xs[q]], v);
    }
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const int sh = pw[q + 1] - pw[q];
      // dir 0: s = t + sh; n = x_q(t) - 1 = x_q(s), m = x_{q+1}(t) = x_{q+1}(s) - 1
      if (xs[q] <= 1 && xs[q + 1] >= 1) lc_cfma(acc[s - sh], cJ[((q * 2 + 0) * 2 + xs[q]) * 2 + xs[q + 1] - 1], v);
      // dir 1: s = t - sh; n = x_q(t) = x_q(s) - 1, m = x_{q+1}(t) - 1 = x_{q+1}(s)
      if (xs[q] >= 1 && xs[q + 1] <= 1) lc_cfma(acc[s + sh], cJ[((q * 2 + 1) * 2 + xs[q] - 1) * 2 + xs[q + 1]], v);
    }
  }
}

// Gather terms.  Every load is unconditional (lanes 27..31 alias lane 26; per-lane invalid columns carry a zero
// coefficient and read initialised memory: the buffers are zero-filled with slack on both sides).
//   * "all rows" terms (operators whose source lies in another row tile with the same in-tile row: X_q, J_q,q+1 and the
//     two-sided terms of sites q >= 3): a run-time list per task, one warp-uniform coefficient each;
//   * terms that depend on an in-tile digit (J_23, the two-sided terms of sites 0..2): straight-line blocks below.
template <int LD>
__device__ __forceinline__ void lc_apply_all(double2 (&acc)[LC_T], const double2* __restrict__ src, const double2 c0) {
#pragma unroll
  for (int t = 0; t < LC_T; t++) lc_cfma(acc[t], c0, __ldg(src + t * LD));
}
// rows whose digit Q is LO take c0, rows whose digit Q is LO + 1 take c1
template <int LD, int Q, int LO>
__device__ __forceinline__ void lc_apply_digit(double2 (&acc)[LC_T], const double2* __restrict__ src, const double2 c0, const double2 c1) {
#pragma unroll
  for (int t = 0; t < LC_T; t++) {
    constexpr int P3 = Q == 0 ? 1 : (Q == 1 ? 3 : 9);
    const int x = (t / P3) % 3;
    if (x == LO) lc_cfma(acc[t], c0, __ldg(src + t * LD));
    if (x == LO + 1) lc_cfma(acc[t], c1, __ldg(src + t * LD));
  }
}

struct LcTask {  // per-warp task description in shared memory (written by lane 0)
  int n_all[3];        // end of the all-rows term list of seg 0, 1, 2 (lists are stored back to back)
  int j23[2][2];       // [row tile: 0 = J (seg 0), 1 = I (seg 1)][dir]: J_23 term present
  int j23_slot[2][2];  // slot of the x_2 = LO coefficient (the LO + 1 one is 2 slots further)
  int off[LC_MAXTERMS];
  int slot[LC_MAXTERMS];
};

template <int NQ, int LD>
__device__ __forceinline__ void lc_build_one_sided(LcTask* tk, int which, int& n, int R, unsigned drv_mask) {
  const int jbase = 4 * NQ;
  int xd[LC_MAXQ + 1];
  int rr = R;
  for (int q = 3; q <= LC_MAXQ; q++) {  // tile digits: x_3 = R % 3, x_4 = (R / 3) % 3, ...
    xd[q] = rr % 3;
    rr /= 3;
  }
  tk->j23[which][0] = tk->j23[which][1] = 0;
  if (NQ >= 4) {  // J_23: digit 2 lives inside the tile, digit 3 in the tile index
    const int x3 = xd[3];
    // dir 0: source r + 27 - 9, rows with x_2(t) = 1, 2: n = x_2(t) - 1, m = x_3.  slot(n, m) = jbase + ((2*2+0)*2 + n)*2 + m
    tk->j23[which][0] = x3 <= 1;
    tk->j23_slot[which][0] = jbase + ((2 * 2 + 0) * 2 + 0) * 2 + x3;
    // dir 1: source r - 27 + 9, rows with x_2(t) = 0, 1: n = x_2(t), m = x_3 - 1
    tk->j23[which][1] = x3 >= 1;
    tk->j23_slot[which][1] = jbase + ((2 * 2 + 1) * 2 + 0) * 2 + x3 - 1;
  }
  int pw = 27;
  for (int q = 3; q < NQ; q++, pw *= 3) {
    const int xq = xd[q];
    if ((drv_mask >> q) & 1u) {
      if (xq <= 1) { tk->off[n] = pw * LD; tk->slot[n++] = (q * 2 + 0) * 2 + xq; }
      if (xq >= 1) { tk->off[n] = -pw * LD; tk->slot[n++] = (q * 2 + 1) * 2 + xq - 1; }
    }
    if (q + 1 < NQ) {
      const int xq1 = xd[q + 1];
      if (xq >= 1 && xq1 <= 1) { tk->off[n] = 2 * pw * LD; tk->slot[n++] = jbase + ((q * 2 + 0) * 2 + xq - 1) * 2 + xq1; }
      if (xq <= 1 && xq1 >= 1) { tk->off[n] = -2 * pw * LD; tk->slot[n++] = jbase + ((q * 2 + 1) * 2 + xq) * 2 + xq1 - 1; }
    }
  }
}

static inline size_t lindblad_chain_smem(int NC) {
  return (size_t)LC_WARPS * ((size_t)LC_T * LC_T * 16 + (size_t)((NC + 1) & ~1) * 16 + ((sizeof(LcTask) + 15) & ~(size_t)15));
}

template <int NQ, int XLOW>
__global__ void __launch_bounds__(32 * LC_WARPS, 1) lindblad_chain_stage_kernel(const LcArgs a, int set, const LbStage s) {
  constexpr int D = lc_pow3(NQ), LD = (D + 7) & ~7;
  extern __shared__ __align__(16) unsigned char lc_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int NCs = (a.NC + 1) & ~1;
  const size_t per_warp = (size_t)LC_T * LC_T * 16 + (size_t)NCs * 16 + ((sizeof(LcTask) + 15) & ~(size_t)15);
  unsigned char* mine = lc_smem + (size_t)warp * per_warp;
  double2* sT = reinterpret_cast<double2*>(mine);              // [27][27]
  double2* ct = sT + LC_T * LC_T;                              // [NC]
  LcTask* tk = reinterpret_cast<LcTask*>(ct + NCs);
  const long long task = (long long)blockIdx.x * LC_WARPS + warp;
  if (task >= a.B * a.npairs) return;
  const long long b = task / a.npairs;
  const int pr = (int)(task - b * a.npairs);
  const int I = a.pair_ij[2 * pr], J = a.pair_ij[2 * pr + 1];
  const bool diag = I == J;
  const bool active = lane < LC_T;
  const int cl = active ? lane : LC_T - 1;  // lanes 27..31 shadow lane 26 (loads stay in bounds, nothing is stored)
  {
    const double2* src = a.ctab + ((long long)set * a.B + b) * a.NC;
    for (int i = lane; i < a.NC; i += 32) ct[i] = __ldg(src + i);
  }
  if (lane == 0) {
    int n = 0;
    if (!diag) lc_build_one_sided<NQ, LD>(tk, 0, n, J, a.drv_mask);
    tk->n_all[0] = n;
    lc_build_one_sided<NQ, LD>(tk, 1, n, I, a.drv_mask);
    tk->n_all[1] = n;
    for (int k = 0; k < a.n2; k++) {  // two-sided terms of the sites in the tile INDEX: warp-uniform coefficient W(k, nr, nc)
      const int q = a.two_q[k];
      if (q <= 2) continue;
      int pw = 1, ti = I, tj = J;
      for (int sq = 3; sq < q; sq++) {
        pw *= 3;
        ti /= 3;
        tj /= 3;
      }
      const int nr = ti % 3, nc = tj % 3;
      if (nr > 1 || nc > 1) continue;
      tk->off[n] = pw * 27 * (LD + 1);
      tk->slot[n++] = a.w_base + k * 4 + nr * 2 + nc;
    }
    tk->n_all[2] = n;
  }
  __syncwarp();
  const long long base = b * (long long)D * LD;
  const double2* in = s.in + base;
  const long long o0 = (long long)(LC_T * I) * LD + LC_T * J + cl;
  const int x0l = cl % 3, x1l = (cl / 3) % 3, x2l = cl / 9;
  double2 acc[LC_T];
  // seg 0: M_JI = (A rho)[J, I] -> sT;  seg 1: M_IJ (diagonal pair: also -> sT);  seg 2: + two-sided terms of (I, J)
#pragma unroll 1
  for (int seg = diag ? 1 : 0; seg < 3; seg++) {
    const double2* p = seg == 0 ? in + (long long)(LC_T * J) * LD + LC_T * I + cl : in + o0;
    if (seg < 2) {
#pragma unroll
      for (int t = 0; t < LC_T; t++) acc[t] = make_double2(0.0, 0.0);
      lc_in_tile<NQ, LD, XLOW>(acc, p, ct);
      if (NQ >= 4) {
        if (tk->j23[seg][0]) {
          const int sl = tk->j23_slot[seg][0];
          lc_apply_digit<LD, 2, 1>(acc, p + 18 * LD, ct[sl], ct[sl + 2]);
        }
        if (tk->j23[seg][1]) {
          const int sl = tk->j23_slot[seg][1];
          lc_apply_digit<LD, 2, 0>(acc, p - 18 * LD, ct[sl], ct[sl + 2]);
        }
      }
    } else {
      // two-sided terms of the sites inside the tile (lowering operators): rho[r + 3^q, c + 3^q], rows and lanes with digit
      // q <= 1; coefficient W(k, nr = x_q(row), nc = x_q(lane))
      const double2 z = make_double2(0.0, 0.0);
      for (int k = 0; k < a.n2; k++) {
        const int q = a.two_q[k];
        if (q > 2) continue;
        const int wb = a.w_base + k * 4;
        const int xl = q == 0 ? x0l : (q == 1 ? x1l : x2l);
        const bool ok = xl <= 1;
        const double2 c0 = ok ? ct[wb + xl] : z, c1 = ok ? ct[wb + 2 + xl] : z;
        if (q == 0) lc_apply_digit<LD, 0, 0>(acc, p + (LD + 1), c0, c1);
        if (q == 1) lc_apply_digit<LD, 1, 0>(acc, p + 3 * (LD + 1), c0, c1);
        if (q == 2) lc_apply_digit<LD, 2, 0>(acc, p + 9 * (LD + 1), c0, c1);
      }
    }
    {
      const int first = seg == 0 ? 0 : tk->n_all[seg - 1], last = tk->n_all[seg];
#pragma unroll 1
      for (int i = first; i < last; i++) lc_apply_all<LD>(acc, p + tk->off[i], ct[tk->slot[i]]);
    }
    if ((seg == 0 || (seg == 1 && diag)) && active) {
#pragma unroll
      for (int t = 0; t < LC_T; t++) sT[t * LC_T + lane] = acc[t];
    }
  }
  __syncwarp();
  // ---- epilogue: k = M_IJ + M_JI^+ + G o rho, RK4 combination, store; the result replaces the consumed sT slot
  if (active) {
    const double2* p = in + o0;
    const double2* g = a.G + o0;
#pragma unroll
    for (int t = 0; t < LC_T; t++) {
      const long long o = base + o0 + (long long)t * LD;
      const double2 y0 = s.y[o];
      const double2 v = __ldg(p + t * LD), gg = __ldg(g + t * LD);
      const double2 m = sT[lane * LC_T + t];  // M_JI(J*27 + lane, I*27 + t)
      double2 k = make_double2(acc[t].x + m.x, acc[t].y - m.y);
      lc_cfma(k, gg, v);
      const double2 res = lb_rk4_combine(s, o, y0, k.x, k.y);
      s.out[o] = res;
      if (!diag) sT[lane * LC_T + t] = res;
    }
  }
  if (diag) return;
  __syncwarp();
  // ---- mirror tile: out(J*27 + t, I*27 + lane) = conj(out(I*27 + lane, J*27 + t)) = conj(sT[t][lane])
  if (active) {
    const long long om = base + (long long)(LC_T * J) * LD + LC_T * I + lane;
#pragma unroll
    for (int t = 0; t < LC_T; t++) {
      const double2 r = sT[t * LC_T + lane];
      s.out[om + (long long)t * LD] = make_double2(r.x, -r.y);
    }
  }
}
