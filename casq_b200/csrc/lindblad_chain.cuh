// casq_b200: Lindblad RK4 stage kernel for CHAINS OF THREE-LEVEL SITES (BASELINE cfg 4: the five-transmon ibmq_manila model,
// d = 3^5 = 243).  casq's TransmonModel always builds this structure: three levels per qubit (src/casq/models/
// transmon_model.py:150), single-site ladder drives X_q, nearest-neighbour exchange couplings, and the T1/T2 noise model
// adds single-site lowering operators and diagonal dephasing operators.  lindblad.cu detects the structure from the dense
// operators handed over the C ABI (every element must fit; anything else runs the generic shift-class kernel of
// lindblad_stage.cuh) and this kernel exploits it:
//
//   index r = sum_q x_q 3^q.   One BLOCK of three warps owns the 27 x 27 tile pair (I,J), (J,I), I <= J, of one trajectory;
//   warp g owns the nine rows with x_2 = g, lane = column inside the tile, and a lane keeps the 9 accumulators of its
//   column (digits x_0, x_1 of the row index) in REGISTERS.  The operators that act inside the low two sites (X_0, X_1,
//   J_01) are applied from ONE load of each source element; everything else is a "gather term" with a warp-uniform
//   coefficient (x_2 and the tile digits are warp-uniform: X_2.., J_23, J_34..) or a short straight-line block (J_12, the
//   two-sided terms L rho L^+ of sites 0..2 with per-lane coefficients).  Against the generic kernel: no per-row lists, no
//   address arithmetic (all offsets are immediates), ~35 % fewer gathered rows, and ~20 resident warps per SM with nine
//   independent loads in flight each (a first version with 27 accumulators per lane needed 252 registers: 8 warps per
//   SM, 57 % of its stall samples on the load scoreboard, 212 traj/s).
//   k = M + M^+ + sum_j L_j rho L_j^+ with M = A rho is formed as in the generic kernel: M_JI is computed with left gathers,
//   transposed through a per-warp shared tile, the RK4 combination is fused, and the mirror tile out[J,I] = out[I,J]^+ is
//   written from the shared tile, so every stage output is exactly Hermitian and y[J,I], y1[J,I].. are never read.
//   Coefficients (value x time function) are a 1 kB table per (trajectory, stage time) written once per step by
//   lindblad_chain_coef_kernel -- the generic path's per-row lists (tens of MB per step) are gone.
#pragma once
#include "lindblad_stage.cuh"

#define LC_T 27          // tile edge = 3 sites
#ifndef LC_TS
#define LC_TS 27         // storage stride of a tile row.  27 = no padding: bulk copies need 16 B alignment only, and padding to 32
#endif                   // (128 B aligned rows, a gain for neither kernel) costs 18 % more HBM / L2 / shared-memory bytes
#define LC_GS ((LC_TS + 1) & ~1)  // row stride of the REAL diagonal-coefficient table (doubles): keeps every nine-row group 16 B aligned
#define LC_WARPS 3       // warps per block = per tile pair: warp g owns the rows with x_2 = g
#define LC_MAXQ 6        // sites
#ifndef LC_MINBLOCKS
#define LC_MINBLOCKS 5    // resident blocks per SM the register allocation aims at: 15 warps, 128 registers, no spills
                         // (4 / 5 / 6 / 7 blocks measured within 2 % of each other: the stage is not occupancy bound)
#endif
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
  const double* Gre;          // the same as reals when every imaginary part is zero (frame = diag(H0)): half the bytes; else NULL
  // coefficient kernel inputs: slot s = sum over atoms [slot_start[s], slot_start[s+1]) of atom_val * tf[atom_code]
  const int* slot_start;
  const double2* atom_val;
  const int* atom_code;
  const void* plan;           // [npairs][3 warps][3 segments] LcList: the gather lists of every task, built once per solve
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

__device__ __forceinline__ void lc_zero9(double2 (&acc)[9]) {
#pragma unroll
  for (int t = 0; t < 9; t++) acc[t] = make_double2(0.0, 0.0);
}

// nine rows of one group from another place (same local rows), one coefficient
template <int LD>
__device__ __forceinline__ void lc_apply9(double2 (&acc)[9], const double2* __restrict__ src, const double2 c) {
  double2 v[9];
#pragma unroll
  for (int t = 0; t < 9; t++) v[t] = __ldg(src + t * LD);
#pragma unroll
  for (int t = 0; t < 9; t++) lc_cfma(acc[t], c, v[t]);
}

// rows of the group whose digit Q (0 or 1) is LO take c0, those whose digit is LO + 1 take c1 (six loads)
template <int LD, int Q, int LO>
__device__ __forceinline__ void lc_apply_digit9(double2 (&acc)[9], const double2* __restrict__ src, const double2 c0, const double2 c1) {
  double2 v[9];
#pragma unroll
  for (int t = 0; t < 9; t++) {
    const int x = Q == 0 ? t % 3 : t / 3;
    if (x == LO || x == LO + 1) v[t] = __ldg(src + t * LD);
  }
#pragma unroll
  for (int t = 0; t < 9; t++) {
    const int x = Q == 0 ? t % 3 : t / 3;
    if (x == LO) lc_cfma(acc[t], c0, v[t]);
    if (x == LO + 1) lc_cfma(acc[t], c1, v[t]);
  }
}

// The part of M = A rho that stays inside the nine rows (x_0, x_1) of a group: X_0, X_1 ladder and J_01 exchange elements,
// every source row loaded once.  XLOW bit q = the ladder slots of site q are compiled in.
template <int NQ, int LD, int XLOW>
__device__ __forceinline__ void lc_in_group(double2 (&acc)[9], const double2* __restrict__ p, const double2* ct) {
  const double2* cJ = ct + 4 * NQ;
  double2 v[9];
#pragma unroll
  for (int s = 0; s < 9; s++) v[s] = __ldg(p + s * LD);
#pragma unroll
  for (int s = 0; s < 9; s++) {
    const int x0 = s % 3, x1 = s / 3;
    if (XLOW & 1) {
      if (x0 >= 1) lc_cfma(acc[s - 1], ct[(0 * 2 + 0) * 2 + x0 - 1], v[s]);  // A[t, s = t + 1]: X(0, 0, n = x_0(t))
      if (x0 <= 1) lc_cfma(acc[s + 1], ct[(0 * 2 + 1) * 2 + x0], v[s]);      // A[t, s = t - 1]: X(0, 1, n = x_0(t) - 1)
    }
    if (XLOW & 2) {
      if (x1 >= 1) lc_cfma(acc[s - 3], ct[(1 * 2 + 0) * 2 + x1 - 1], v[s]);
      if (x1 <= 1) lc_cfma(acc[s + 3], ct[(1 * 2 + 1) * 2 + x1], v[s]);
    }
    // J_01, s = t + 2: n = x_0(t) - 1 = x_0(s), m = x_1(t) = x_1(s) - 1;  s = t - 2: n = x_0(t) = x_0(s) - 1, m = x_1(t) - 1 = x_1(s)
    if (x0 <= 1 && x1 >= 1) lc_cfma(acc[s - 2], cJ[((0 * 2 + 0) * 2 + x0) * 2 + x1 - 1], v[s]);
    if (x0 >= 1 && x1 <= 1) lc_cfma(acc[s + 2], cJ[((0 * 2 + 1) * 2 + x0 - 1) * 2 + x1], v[s]);
  }
}

struct LcList {  // gather terms of one warp for one segment: nine rows each, warp-uniform coefficient
  int n;
  int off[15];
  int slot[15];
  int pad;
};

// one-sided gather terms of the rows (R, x_2 = g): everything of A that leaves the nine-row group except J_12.
// Offsets in elements of the tile-major storage: RS = row stride inside a tile, TROW = stride between tile rows.
template <int NQ, int RS, int TROW>
__device__ __forceinline__ void lc_build_one_sided(LcList* L, int R, int g, unsigned drv_mask) {
  const int jbase = 4 * NQ;
  int n = 0;
  int xd[LC_MAXQ + 1];
  int rr = R;
  for (int q = 3; q <= LC_MAXQ; q++) {  // tile digits: x_3 = R % 3, x_4 = (R / 3) % 3, ...
    xd[q] = rr % 3;
    rr /= 3;
  }
  if ((drv_mask >> 2) & 1u) {  // X_2: source group g + 1 / g - 1 of the same tile
    if (g <= 1) { L->off[n] = 9 * RS; L->slot[n++] = (2 * 2 + 0) * 2 + g; }
    if (g >= 1) { L->off[n] = -9 * RS; L->slot[n++] = (2 * 2 + 1) * 2 + g - 1; }
  }
  if (NQ >= 4) {  // J_23: r + 27 - 9 (x_2 >= 1, x_3 <= 1: n = x_2 - 1, m = x_3) and r - 27 + 9 (x_2 <= 1, x_3 >= 1: n = x_2, m = x_3 - 1)
    const int x3 = xd[3];
    if (g >= 1 && x3 <= 1) { L->off[n] = TROW - 9 * RS; L->slot[n++] = jbase + ((2 * 2 + 0) * 2 + g - 1) * 2 + x3; }
    if (g <= 1 && x3 >= 1) { L->off[n] = -(TROW - 9 * RS); L->slot[n++] = jbase + ((2 * 2 + 1) * 2 + g) * 2 + x3 - 1; }
  }
  int tw = 1;  // 3^(q - 3): shift in tile rows
  for (int q = 3; q < NQ; q++, tw *= 3) {
    const int xq = xd[q];
    if ((drv_mask >> q) & 1u) {
      if (xq <= 1) { L->off[n] = tw * TROW; L->slot[n++] = (q * 2 + 0) * 2 + xq; }
      if (xq >= 1) { L->off[n] = -tw * TROW; L->slot[n++] = (q * 2 + 1) * 2 + xq - 1; }
    }
    if (q + 1 < NQ) {
      const int xq1 = xd[q + 1];
      if (xq >= 1 && xq1 <= 1) { L->off[n] = 2 * tw * TROW; L->slot[n++] = jbase + ((q * 2 + 0) * 2 + xq - 1) * 2 + xq1; }
      if (xq <= 1 && xq1 >= 1) { L->off[n] = -2 * tw * TROW; L->slot[n++] = jbase + ((q * 2 + 1) * 2 + xq) * 2 + xq1 - 1; }
    }
  }
  L->n = n;
}

static inline size_t lindblad_chain_smem(int NC) {
  return (size_t)LC_T * LC_T * 16 + (size_t)((NC + 1) & ~1) * 16 + 3 * LC_WARPS * sizeof(LcList);
}

template <int NQ, int XLOW>
__global__ void __launch_bounds__(32 * LC_WARPS, LC_MINBLOCKS) lindblad_chain_stage_kernel(const LcArgs a, int set, const LbStage s) {
  constexpr int D = lc_pow3(NQ), NT = D / LC_T, LD = NT * LC_TS, RS = LC_TS, TSZ = LC_T * LC_TS, TROW = NT * TSZ;  // tile-major (lb_idx)
  extern __shared__ __align__(16) unsigned char lc_smem[];
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int NCs = (a.NC + 1) & ~1;
  double2* sT = reinterpret_cast<double2*>(lc_smem);           // [27][27]
  double2* ct = sT + LC_T * LC_T;                              // [NC]
  LcList* lists = reinterpret_cast<LcList*>(ct + NCs) + 3 * g; // [warp][seg]
  const long long task = blockIdx.x;
  const long long b = task / a.npairs;
  const int pr = (int)(task - b * a.npairs);
  const int I = a.pair_ij[2 * pr], J = a.pair_ij[2 * pr + 1];
  const bool diag = I == J;
  const bool active = lane < LC_T;
  const int cl = active ? lane : LC_T - 1;  // lanes 27..31 shadow lane 26 (loads stay in bounds, nothing is stored)
  {
    const double2* src = a.ctab + ((long long)set * a.B + b) * a.NC;
    for (int i = threadIdx.x; i < a.NC; i += 32 * LC_WARPS) ct[i] = __ldg(src + i);
  }
  if (lane == 0) {
    if (!diag) lc_build_one_sided<NQ, RS, TROW>(lists + 0, J, g, a.drv_mask);
    lc_build_one_sided<NQ, RS, TROW>(lists + 1, I, g, a.drv_mask);
    int n = 0;
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
      lists[2].off[n] = pw * (TROW + TSZ);
      lists[2].slot[n++] = a.w_base + k * 4 + nr * 2 + nc;
    }
    lists[2].n = n;
  }
  __syncthreads();
  const long long base = b * (long long)D * LD;
  const double2* in = s.in + base;
  const long long o0 = ((long long)I * NT + J) * TSZ + 9 * g * RS + cl;   // this warp's rows of tile (I, J)
  const double2* cJ12 = ct + 4 * NQ + (1 * 2 + 0) * 4;  // J(1, dir, n, m) = cJ12[(dir*2 + n)*2 + m]
  double2 acc[9];
  // seg 0: M_JI = (A rho)[J, I] -> sT;  seg 1: M_IJ (diagonal pair: also -> sT);  seg 2: + two-sided terms of (I, J)
#pragma unroll 1
  for (int seg = diag ? 1 : 0; seg < 3; seg++) {
    const double2* p = seg == 0 ? in + ((long long)J * NT + I) * TSZ + 9 * g * RS + cl : in + o0;
    if (seg < 2) {
      lc_zero9(acc);
      lc_in_group<NQ, RS, XLOW>(acc, p, ct);
      // J_12 inside the tile: s = t + 6 (x_1(t) >= 1, g <= 1: n = x_1(t) - 1, m = g);  s = t - 6 (x_1(t) <= 1, g >= 1: n = x_1(t), m = g - 1)
      if (g <= 1) lc_apply_digit9<RS, 1, 1>(acc, p + 6 * RS, cJ12[(0 * 2 + 0) * 2 + g], cJ12[(0 * 2 + 1) * 2 + g]);
      if (g >= 1) lc_apply_digit9<RS, 1, 0>(acc, p - 6 * RS, cJ12[(1 * 2 + 0) * 2 + g - 1], cJ12[(1 * 2 + 1) * 2 + g - 1]);
    } else {
      // two-sided terms of the sites inside the tile (lowering operators): rho[r + 3^q, c + 3^q] for rows and lanes whose
      // digit q is <= 1; coefficient W(k, nr = x_q(row), nc = x_q(lane)) = ct[w_base + 4k + 2 nr + nc]
      const double2 z = make_double2(0.0, 0.0);
      for (int k = 0; k < a.n2; k++) {
        const int q = a.two_q[k];
        if (q > 2) continue;
        const int wb = a.w_base + k * 4;
        const int xl = q == 0 ? cl % 3 : (q == 1 ? (cl / 3) % 3 : cl / 9);
        const bool ok = xl <= 1;
        const int xx = ok ? xl : 0;
        if (q == 0) lc_apply_digit9<RS, 0, 0>(acc, p + (RS + 1), ok ? ct[wb + xx] : z, ok ? ct[wb + 2 + xx] : z);
        if (q == 1) lc_apply_digit9<RS, 1, 0>(acc, p + 3 * (RS + 1), ok ? ct[wb + xx] : z, ok ? ct[wb + 2 + xx] : z);
        if (q == 2 && g <= 1) lc_apply9<RS>(acc, p + 9 * (RS + 1), ok ? ct[wb + 2 * g + xx] : z);
      }
    }
    {
      const LcList* L = lists + seg;
      const int n = L->n;
#pragma unroll 1
      for (int i = 0; i < n; i++) lc_apply9<RS>(acc, p + L->off[i], ct[L->slot[i]]);
    }
    if ((seg == 0 || (seg == 1 && diag)) && active) {
#pragma unroll
      for (int t = 0; t < 9; t++) sT[(9 * g + t) * LC_T + lane] = acc[t];
    }
  }
  __syncthreads();
  // ---- epilogue: k = M_IJ + M_JI^+ + G o rho, RK4 combination, store; the result replaces the consumed sT slot
  if (active) {
    const double2* p = in + o0;
    const long long ob = base + o0;
    const double2* yp = s.y + ob;
    const double2* y1p = s.y1 + ob;
    const double2* y2p = s.y2 + ob;
    double2* outp = s.out + ob;
    const bool greal = a.Gre != nullptr;
    const double2* gp = a.G + o0;
    const double* grp = greal ? a.Gre + ((long long)I * NT + J) * (LC_T * LC_GS) + 9 * g * LC_GS + cl : nullptr;
    const double h = s.stage == 3 ? s.h : 0.5 * s.h, t3 = 1.0 / 3.0, h6 = s.h * (1.0 / 6.0);
#pragma unroll
    for (int t0 = 0; t0 < 9; t0 += 3) {  // three rows at a time: nine or more independent loads in flight
      double2 y0[3], v[3], gg[3];
#pragma unroll
      for (int u = 0; u < 3; u++) {
        y0[u] = yp[(t0 + u) * RS];
        v[u] = __ldg(p + (t0 + u) * RS);
        gg[u] = greal ? make_double2(__ldg(grp + (t0 + u) * LC_GS), 0.0) : __ldg(gp + (t0 + u) * RS);
      }
#pragma unroll
      for (int u = 0; u < 3; u++) {
        const int t = t0 + u;
        const double2 m = sT[lane * LC_T + 9 * g + t];  // M_JI(J*27 + lane, I*27 + 9g + t)
        double2 k = make_double2(acc[t].x + m.x, acc[t].y - m.y);
        lc_cfma(k, gg[u], v[u]);
        double2 res;
        if (s.stage != 4) {
          res = make_double2(fma(h, k.x, y0[u].x), fma(h, k.y, y0[u].y));
        } else {  // y <- (y1 + 2 y2 + y3 - y)/3 + h/6 k4 (v = y3 is the stage input)
          const double2 v1 = y1p[t * RS], v2 = y2p[t * RS];
          res = make_double2(fma(h6, k.x, t3 * (v1.x + 2.0 * v2.x + v[u].x - y0[u].x)),
                             fma(h6, k.y, t3 * (v1.y + 2.0 * v2.y + v[u].y - y0[u].y)));
        }
        outp[t * RS] = res;
        if (!diag) sT[lane * LC_T + 9 * g + t] = res;
      }
    }
  }
  if (diag) return;
  __syncthreads();
  // ---- mirror tile: out(J*27 + r, I*27 + lane) = conj(out(I*27 + lane, J*27 + r)) = conj(sT[r][lane]), r = 9g + t
  if (active) {
    const long long om = base + ((long long)J * NT + I) * TSZ + 9 * g * RS + lane;
#pragma unroll
    for (int t = 0; t < 9; t++) {
      const double2 r = sT[(9 * g + t) * LC_T + lane];
      s.out[om + t * RS] = make_double2(r.x, -r.y);
    }
  }
}

// =====================================================================================================================
// TMA-staged variant (default).  Same task decomposition and arithmetic as lindblad_chain_stage_kernel above, but every
// nine-row group a warp consumes is brought from global memory into a per-warp shared-memory ring by bulk asynchronous
// copies (cp.async.bulk -> SASS UBLKCP, completion on an mbarrier: the "coalesced HBM staging of the batch" north_star asks
// for) issued up to LC_NBUF groups ahead of their use:
//   * the LSU no longer carries the loads: a 432 B row read with LDG.128 costs ~8 L1 cycles (2 per 128 B line within one
//     instruction, B300_MICROARCH "rt_L1tex_wf"), the same row read back from shared memory 4;
//   * HBM / L2 latency overlaps the FMA phases without more warps or registers (the LDG version spent 50 % of its stall
//     samples on the load scoreboard at every occupancy from 12 to 21 warps per SM).
// The fetch list of a warp is built in consumption order by its lane 0; lanes 0..8 issue one row copy each.
#ifndef LC_NBUF
#define LC_NBUF 2
#endif
#ifndef LC_TMA_MINBLOCKS
#define LC_TMA_MINBLOCKS 4  // shared memory (50 kB per block with two ring slots per warp) allows four blocks = 12 warps per SM
#endif
#define LC_ROWB (LC_TS * 16)       // bytes of one staged row (a tile row)
#define LC_SLOTB (9 * LC_ROWB)     // one ring slot: nine rows = one contiguous piece of the tile-major storage
#define LC_MAXFETCH 48

struct LcFetch {
  const char* src;   // contiguous in the tile-major storage
  int bytes;         // multiple of 16, <= LC_SLOTB
  int pad;
};

__device__ __forceinline__ unsigned lc_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lc_mbar_init(unsigned mbar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count));
}
__device__ __forceinline__ void lc_mbar_expect_tx(unsigned mbar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void lc_mbar_wait(unsigned mbar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LC_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LC_DONE_%=;\n"
      "bra LC_WAIT_%=;\n"
      "LC_DONE_%=:\n"
      "}\n" ::"r"(mbar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void lc_bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(mbar)
               : "memory");
}

// shared memory of a block: [tile A: (J, I) tile of the stage input, later M_JI^T / the result tile][tile B: (I, J) tile]
// [coefficient table][gather lists][per warp: ring of LC_NBUF slots, fetch list, mbarriers]
#define LC_TILEB (LC_T * LC_ROWB)  // one staged tile
static inline size_t lindblad_chain_tma_smem(int NC) {
  return 2 * (size_t)LC_TILEB + (size_t)((NC + 1) & ~1) * 16 + 3 * LC_WARPS * sizeof(LcList) + 64 +
         LC_WARPS * ((size_t)LC_NBUF * LC_SLOTB + LC_MAXFETCH * sizeof(LcFetch) + 64);
}

// consumers of a staged slot (rows at pitch 27 elements): same arithmetic as lc_in_group / lc_apply9 / lc_apply_digit9
template <int NQ, int XLOW>
__device__ __forceinline__ void lc_s_in_group(double2 (&acc)[9], const double2* __restrict__ b, const double2* ct) {
  const double2* cJ = ct + 4 * NQ;
  double2 v[9];
#pragma unroll
  for (int s = 0; s < 9; s++) v[s] = b[s * LC_TS];
#pragma unroll
  for (int s = 0; s < 9; s++) {
    const int x0 = s % 3, x1 = s / 3;
    if (XLOW & 1) {
      if (x0 >= 1) lc_cfma(acc[s - 1], ct[(0 * 2 + 0) * 2 + x0 - 1], v[s]);
      if (x0 <= 1) lc_cfma(acc[s + 1], ct[(0 * 2 + 1) * 2 + x0], v[s]);
    }
    if (XLOW & 2) {
      if (x1 >= 1) lc_cfma(acc[s - 3], ct[(1 * 2 + 0) * 2 + x1 - 1], v[s]);
      if (x1 <= 1) lc_cfma(acc[s + 3], ct[(1 * 2 + 1) * 2 + x1], v[s]);
    }
    if (x0 <= 1 && x1 >= 1) lc_cfma(acc[s - 2], cJ[((0 * 2 + 0) * 2 + x0) * 2 + x1 - 1], v[s]);
    if (x0 >= 1 && x1 <= 1) lc_cfma(acc[s + 2], cJ[((0 * 2 + 1) * 2 + x0 - 1) * 2 + x1], v[s]);
  }
}
__device__ __forceinline__ void lc_s_apply9(double2 (&acc)[9], const double2* __restrict__ b, const double2 c) {
#pragma unroll
  for (int t = 0; t < 9; t++) lc_cfma(acc[t], c, b[t * LC_TS]);
}
// staged rows 0..5 feed accumulators FIRST..FIRST+5: digit x_1 of the accumulator row picks the coefficient
template <int FIRST>
__device__ __forceinline__ void lc_s_apply6(double2 (&acc)[9], const double2* __restrict__ b, const double2 c0, const double2 c1) {
#pragma unroll
  for (int t = 0; t < 6; t++) lc_cfma(acc[FIRST + t], t < 3 ? c0 : c1, b[t * LC_TS]);
}
// staged window row t feeds accumulator t when digit Q of t is 0 (c0) or 1 (c1)
template <int Q>
__device__ __forceinline__ void lc_s_apply_digit9(double2 (&acc)[9], const double2* __restrict__ b, const double2 c0, const double2 c1) {
#pragma unroll
  for (int t = 0; t < (Q == 0 ? 8 : 6); t++) {  // only these window rows are staged
    const int x = Q == 0 ? t % 3 : t / 3;
    if (x == 0) lc_cfma(acc[t], c0, b[t * LC_TS]);
    if (x == 1) lc_cfma(acc[t], c1, b[t * LC_TS]);
  }
}

// The gather lists depend on the tile pair and the row group only: one launch per solve writes them for every task shape.
template <int NQ>
__global__ void lindblad_chain_plan_kernel(const LcArgs a, LcList* plan) {
  constexpr int D = lc_pow3(NQ), NT = D / LC_T, RS = LC_TS, TSZ = LC_T * LC_TS, TROW = NT * TSZ;
  const int pr = blockIdx.x, g = threadIdx.x;
  if (pr >= a.npairs || g >= LC_WARPS) return;
  const int I = a.pair_ij[2 * pr], J = a.pair_ij[2 * pr + 1];
  LcList* lists = plan + ((size_t)pr * LC_WARPS + g) * 3;
  if (I != J) lc_build_one_sided<NQ, RS, TROW>(lists + 0, J, g, a.drv_mask);
  else lists[0].n = 0;
  lc_build_one_sided<NQ, RS, TROW>(lists + 1, I, g, a.drv_mask);
  int n = 0;
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
    lists[2].off[n] = pw * (TROW + TSZ);
    lists[2].slot[n++] = a.w_base + k * 4 + nr * 2 + nc;
  }
  lists[2].n = n;
}

template <int NQ, int XLOW>
__global__ void __launch_bounds__(32 * LC_WARPS, LC_TMA_MINBLOCKS) lindblad_chain_tma_kernel(const LcArgs a, int set, const LbStage s) {
  constexpr int D = lc_pow3(NQ), NT = D / LC_T, LD = NT * LC_TS, RS = LC_TS, TSZ = LC_T * LC_TS, TROW = NT * TSZ;  // tile-major (lb_idx)
  extern __shared__ __align__(128) unsigned char lc_smem[];
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int NCs = (a.NC + 1) & ~1;
  double2* tileA = reinterpret_cast<double2*>(lc_smem);              // [27][32]: stage input, tile (J, I)
  double2* sT = tileA;                                               // [27][27] once tile A has been consumed
  double2* tileB = tileA + LC_T * LC_TS;                             // [27][32]: stage input, tile (I, J)
  double2* ct = tileB + LC_T * LC_TS;                                // [NC]
  LcList* lists = reinterpret_cast<LcList*>(ct + NCs) + 3 * g;       // [warp][seg]
  unsigned long long* tbar = reinterpret_cast<unsigned long long*>(reinterpret_cast<LcList*>(ct + NCs) + 3 * LC_WARPS);  // [2] tile barriers
  unsigned char* wbase = reinterpret_cast<unsigned char*>(tbar) + 64 + (size_t)g * ((size_t)LC_NBUF * LC_SLOTB + LC_MAXFETCH * sizeof(LcFetch) + 64);
  unsigned char* ring = wbase;                                              // [LC_NBUF][LC_SLOTB]
  LcFetch* fl = reinterpret_cast<LcFetch*>(wbase + (size_t)LC_NBUF * LC_SLOTB);
  unsigned long long* mbars = reinterpret_cast<unsigned long long*>(fl + LC_MAXFETCH);  // [LC_NBUF] + nfetch
  int* nfetch_p = reinterpret_cast<int*>(mbars + LC_NBUF);
  const long long task = blockIdx.x;
  const long long b = task / a.npairs;
  const int pr = (int)(task - b * a.npairs);
  const int I = a.pair_ij[2 * pr], J = a.pair_ij[2 * pr + 1];
  const bool diag = I == J;
  const bool active = lane < LC_T;
  const int cl = active ? lane : LC_T - 1;
  const long long base = b * (long long)D * LD;
  const double2* in = s.in + base;
  const long long tI = ((long long)I * NT + J) * TSZ + 9 * g * RS;  // this warp's rows of tile (I, J), column 0
  const long long tJ = ((long long)J * NT + I) * TSZ + 9 * g * RS;
  const bool greal = a.Gre != nullptr;
  if (threadIdx.x == 0) {
    lc_mbar_init(lc_smem_u32(tbar + 0), LC_WARPS);
    lc_mbar_init(lc_smem_u32(tbar + 1), LC_WARPS);
  }
  if (lane == 0) {
    for (int i = 0; i < LC_NBUF; i++) lc_mbar_init(lc_smem_u32(mbars + i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  {  // this warp's three gather lists (3 x 128 B) from the per-solve plan
    const int4* src = reinterpret_cast<const int4*>(reinterpret_cast<const LcList*>(a.plan) + ((size_t)pr * LC_WARPS + g) * 3);
    if (lane < (int)(3 * sizeof(LcList) / 16)) reinterpret_cast<int4*>(lists)[lane] = __ldg(src + lane);
  }
  __syncwarp();
  if (lane == 0) {
    // the two tiles of the stage input: every warp brings its own nine rows (one contiguous piece each)
    if (!diag) {
      lc_mbar_expect_tx(lc_smem_u32(tbar + 0), 9 * LC_ROWB);
      lc_bulk_g2s(lc_smem_u32(tileA + 9 * g * LC_TS), in + tJ, 9 * LC_ROWB, lc_smem_u32(tbar + 0));
    }
    lc_mbar_expect_tx(lc_smem_u32(tbar + 1), 9 * LC_ROWB);
    lc_bulk_g2s(lc_smem_u32(tileB + 9 * g * LC_TS), in + tI, 9 * LC_ROWB, lc_smem_u32(tbar + 1));
    // ---- ring fetch list, in consumption order: everything that is not in the two staged tiles
    int nf = 0;
    auto add = [&](const void* p, int bytes) {
      fl[nf].src = reinterpret_cast<const char*>(p);
      fl[nf].bytes = bytes;
      nf++;
    };
    for (int seg = diag ? 1 : 0; seg < 2; seg++)
      for (int i = 0; i < lists[seg].n; i++) add(in + (seg == 0 ? tJ : tI) + lists[seg].off[i], 9 * LC_ROWB);
    for (int i = 0; i < lists[2].n; i++) add(in + tI + lists[2].off[i], 9 * LC_ROWB);
    if (s.stage != 1) add(s.y + base + tI, 9 * LC_ROWB);  // stage 1: y is the stage input itself (tile B)
    if (greal) add(a.Gre + ((long long)I * NT + J) * (LC_T * LC_GS) + 9 * g * LC_GS, 9 * LC_GS * 8);
    else add(a.G + tI, 9 * LC_ROWB);
    if (s.stage == 4) {
      add(s.y1 + base + tI, 9 * LC_ROWB);
      add(s.y2 + base + tI, 9 * LC_ROWB);
    }
    *nfetch_p = nf;
  }
  {
    const double2* src = a.ctab + ((long long)set * a.B + b) * a.NC;
    for (int i = threadIdx.x; i < a.NC; i += 32 * LC_WARPS) ct[i] = __ldg(src + i);
  }
  __syncwarp();
  const int nfetch = *nfetch_p;
  auto issue = [&](int k) {  // fill ring slot k % LC_NBUF with fetch k (lane 0)
    if (lane == 0) {
      const LcFetch f = fl[k];
      const int slot = k % LC_NBUF;
      const unsigned mb = lc_smem_u32(mbars + slot);
      lc_mbar_expect_tx(mb, (unsigned)f.bytes);
      lc_bulk_g2s(lc_smem_u32(ring + (size_t)slot * LC_SLOTB), f.src, (unsigned)f.bytes, mb);
    }
  };
  int kc = 0;  // next fetch to consume
#pragma unroll 1
  for (int k = 0; k < LC_NBUF && k < nfetch; k++) issue(k);
  auto acquire = [&]() -> const double2* {
    const int slot = kc % LC_NBUF;
    lc_mbar_wait(lc_smem_u32(mbars + slot), (unsigned)((kc / LC_NBUF) & 1));
    return reinterpret_cast<const double2*>(ring + (size_t)slot * LC_SLOTB) + cl;
  };
  auto release = [&]() {
    __syncwarp();
    if (kc + LC_NBUF < nfetch) issue(kc + LC_NBUF);
    kc++;
  };
  __syncthreads();  // coefficient table and lists of every warp are in place
  const double2* cJ12 = ct + 4 * NQ + (1 * 2 + 0) * 4;
  double2 acc[9];
#pragma unroll 1
  for (int seg = diag ? 1 : 0; seg < 3; seg++) {
    if (seg < 2) {
      lc_mbar_wait(lc_smem_u32(tbar + seg), 0);
      const double2* tp = (seg == 0 ? tileA : tileB) + 9 * g * LC_TS + cl;  // this warp's rows, this lane's column
      lc_zero9(acc);
      lc_s_in_group<NQ, XLOW>(acc, tp, ct);
      // J_12 from the sibling groups of the same staged tile
      if (g <= 1) lc_s_apply6<3>(acc, tp + 9 * LC_TS, cJ12[(0 * 2 + 0) * 2 + g], cJ12[(0 * 2 + 1) * 2 + g]);
      if (g >= 1) lc_s_apply6<0>(acc, tp - 6 * LC_TS, cJ12[(1 * 2 + 0) * 2 + g - 1], cJ12[(1 * 2 + 1) * 2 + g - 1]);
    } else {
      // two-sided terms of the sites inside the tile, from the staged (I, J) tile: rho[r + 3^q, c + 3^q]
      const double2 z = make_double2(0.0, 0.0);
      for (int k = 0; k < a.n2; k++) {
        const int q = a.two_q[k];
        if (q > 2 || (q == 2 && g > 1)) continue;
        const int wb = a.w_base + k * 4;
        const int xl = q == 0 ? cl % 3 : (q == 1 ? (cl / 3) % 3 : cl / 9);
        const bool ok = xl <= 1;
        const int xx = ok ? xl : 0;
        const int sh = q == 0 ? 1 : (q == 1 ? 3 : 9);
        // lanes whose digit is 2 have no term: they read their own column (finite data) with a zero coefficient
        const double2* wp = tileB + (9 * g + sh) * LC_TS + (ok ? cl + sh : cl);
        if (q == 0) lc_s_apply_digit9<0>(acc, wp, ok ? ct[wb + xx] : z, ok ? ct[wb + 2 + xx] : z);
        if (q == 1) lc_s_apply_digit9<1>(acc, wp, ok ? ct[wb + xx] : z, ok ? ct[wb + 2 + xx] : z);
        if (q == 2) lc_s_apply9(acc, wp, ok ? ct[wb + 2 * g + xx] : z);
      }
    }
    {
      const LcList* L = lists + seg;
      const int n = L->n;
#pragma unroll 1
      for (int i = 0; i < n; i++) {
        const double2* bp = acquire();
        lc_s_apply9(acc, bp, ct[L->slot[i]]);
        release();
      }
    }
    if (seg == 0 || (seg == 1 && diag)) {
      if (seg == 0) __syncthreads();  // every warp is done with tile A before M_JI overwrites it
      if (active) {
#pragma unroll
        for (int t = 0; t < 9; t++) sT[(9 * g + t) * LC_T + lane] = acc[t];
      }
    }
  }
  __syncthreads();
  // ---- epilogue
  {
    double2 y0[9];
    const double2* vp = tileB + 9 * g * LC_TS + cl;  // the stage input itself
    if (s.stage != 1) {
      const double2* bp = acquire();
#pragma unroll
      for (int t = 0; t < 9; t++) y0[t] = bp[t * LC_TS];
      release();
    } else {
#pragma unroll
      for (int t = 0; t < 9; t++) y0[t] = vp[t * LC_TS];
    }
    double2 k9[9];
    {
      const double2* bp = acquire();
#pragma unroll
      for (int t = 0; t < 9; t++) {
        double2 gg;
        if (greal) gg = make_double2(reinterpret_cast<const double*>(bp - cl)[t * LC_GS + cl], 0.0);
        else gg = bp[t * LC_TS];
        const double2 m = sT[cl * LC_T + 9 * g + t];  // M_JI(J*27 + lane, I*27 + 9g + t)
        double2 k = make_double2(acc[t].x + m.x, acc[t].y - m.y);
        lc_cfma(k, gg, vp[t * LC_TS]);
        k9[t] = k;
      }
      release();
    }
    const double h = s.stage == 3 ? s.h : 0.5 * s.h, t3 = 1.0 / 3.0, h6 = s.h * (1.0 / 6.0);
    double2 res[9];
    if (s.stage != 4) {
#pragma unroll
      for (int t = 0; t < 9; t++) res[t] = make_double2(fma(h, k9[t].x, y0[t].x), fma(h, k9[t].y, y0[t].y));
    } else {
      const double2* b1 = acquire();
      double2 w[9];
#pragma unroll
      for (int t = 0; t < 9; t++) w[t] = b1[t * LC_TS];
      release();
      const double2* b2 = acquire();
#pragma unroll
      for (int t = 0; t < 9; t++) {
        const double2 v2 = b2[t * LC_TS], v3 = vp[t * LC_TS];
        res[t] = make_double2(fma(h6, k9[t].x, t3 * (w[t].x + 2.0 * v2.x + v3.x - y0[t].x)),
                              fma(h6, k9[t].y, t3 * (w[t].y + 2.0 * v2.y + v3.y - y0[t].y)));
      }
      release();
    }
    if (active) {
      double2* outp = s.out + base + tI + lane;
#pragma unroll
      for (int t = 0; t < 9; t++) {
        outp[t * RS] = res[t];
        if (!diag) sT[lane * LC_T + 9 * g + t] = res[t];
      }
    }
  }
  if (diag) return;
  __syncthreads();
  if (active) {
    const long long om = base + tJ + lane;
#pragma unroll
    for (int t = 0; t < 9; t++) {
      const double2 r = sT[(9 * g + t) * LC_T + lane];
      s.out[om + t * RS] = make_double2(r.x, -r.y);
    }
  }
}
