// casq_b200: device side of the structured Lindblad RK4 path (see lindblad.cu for the host side and the maths).
//
// Per RK4 stage:   k = Lm + Lm^+ + sum_j L_j rho L_j^+      with   Lm = A(t) rho,   A = -i H_F(t) - 1/2 D_F(t)
// where A is a SPARSE d x d matrix (local operators on a product basis: ~6 non-zeros per row for ibmq_manila)
// whose elements are value x time-function look-ups ("shift classes", lindblad.cu).
//
//   lindblad_coef_kernel     once per step: per trajectory and stage time, the class coefficient vectors and, per
//                            row, the COMPACTED list of (coefficient, source-row offset) of the classes acting on it.
//   lindblad_spmm_kernel     Lm = A rho as a row-wise SpMM: one warp owns a full output row, lanes own columns
//                            (8 x 32 column chunks in registers), so each (coefficient, offset) pair is loaded once
//                            per row and amortised over 243 elements: the inner loop is LDG.128 + 4 DFMA.
//   lindblad_combine_kernel  k = Lm + Lm^+ (transposed tile through shared memory) + two-sided gathers, then the
//                            fused RK4 combination with four rolling buffers.
//
// Round-1 history (profiles/r01_ncu_lindblad_*.txt): tile-pair kernels that gathered per 27x27 tile spent
// >90 % of their instructions on per-tile set-up and index arithmetic (1 500 thread-instructions per element
// pair against ~100 useful DFMA) and ran at the same 75-85 traj/s whatever the gather count or block shape.
#pragma once
#include "common.cuh"
#include "envelope.cuh"

#define LB_MAXK 4
#define LB_MAXQ 32  // one-sided classes (list slots per row)

struct LbArgs {
  int d;
  long long B;
  int KA, NF, NTF, NQ, NU, NG, NP, W;
  const int* q_delta;     // [NQ]
  const short* q_code;    // [NQ][d]  time-function id or -1
  const double2* q_val;   // [NQ][d]
  const int* u_delta;     // [NU]
  const short* u_code;    // [NU][d]
  const double2* u_val;   // [NU][d]
  const int* grp_start;   // [NG+1] two-sided class pairs grouped by (delta1, delta2): they share one gathered element
  const int* grp_pairs;   // [NP][2]
  const int* tf_sig;      // [NTF] 0 const, 1+k: z_k, 1+KA+k: conj z_k, 1+2KA+k: Re z_k
  const int* tf_freq;     // [NTF] index into the phase part of the stage table, -1 = no phase
  const double* table;    // [NT][W]: KA carriers (cos,sin) then NF phases (cos,sin)
  const int* idx;         // [NT] sample index
  const double* params;
  DevSlots slots;
  int rows_per_block;     // SpMM: consecutive rows walked by one block
  const double2* cdense;  // [d*d] constant coefficient of the two-sided terms that gather rho[r,c] itself (x 1/2), or NULL
};

struct LbLists {   // per (set, trajectory, row), written by lindblad_coef_kernel
  int* lcnt;       // [3][B][d]       number of one-sided classes acting on the row
  double2* lcoef;  // [3][B][d][NQ]   their coefficients
  int* loff;       // [3][B][d][NQ]   element offset delta*d of the source row
  int* gcnt;       // [3][B][d]       number of two-sided gather groups with a non-zero row factor
  int* glist;      // [3][B][d][NG]
};

__device__ __forceinline__ double2 lb_cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

__global__ void __launch_bounds__(256) lindblad_coef_kernel(const LbArgs a, long long entry0, double2* __restrict__ cq,
                                                            double2* __restrict__ cu, const LbLists L) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* s_tf = reinterpret_cast<double2*>(smem_raw);  // [NTF]
  __shared__ dcplx s_env[LB_MAXK];
  const long long b = blockIdx.x;
  const int set = blockIdx.y, d = a.d, tid = threadIdx.x;
  const long long entry = entry0 + set;
  if (tid == 0) {
    dcplx env[LB_MAXK];
    casq_channel_envelopes<LB_MAXK>(a.slots, a.params, a.B, b, __ldg(a.idx + entry), env);
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
      const double c = ent[2 * k], s = ent[2 * k + 1];
      const double zr = s_env[k].re * c - s_env[k].im * s, zi = s_env[k].re * s + s_env[k].im * c;
      v = kind == 0 ? make_double2(zr, zi) : (kind == 1 ? make_double2(zr, -zi) : make_double2(zr, 0.0));
    }
    if (f >= 0) v = lb_cmul(v, make_double2(ent[2 * a.KA + 2 * f], ent[2 * a.KA + 2 * f + 1]));
    s_tf[n] = v;
  }
  __syncthreads();
  double2* oq = cq + ((long long)set * a.B + b) * a.NQ * d;
  for (int e = tid; e < a.NQ * d; e += blockDim.x) {
    const int code = a.q_code[e];
    oq[e] = code >= 0 ? lb_cmul(a.q_val[e], s_tf[code]) : make_double2(0.0, 0.0);
  }
  double2* ou = cu + ((long long)set * a.B + b) * a.NU * d;
  for (int e = tid; e < a.NU * d; e += blockDim.x) {
    const int code = a.u_code[e];
    ou[e] = code >= 0 ? lb_cmul(a.u_val[e], s_tf[code]) : make_double2(0.0, 0.0);
  }
  __syncthreads();  // this block wrote oq / ou itself
  const long long rowbase = ((long long)set * a.B + b) * d;
  for (int row = tid; row < d; row += blockDim.x) {
    double2* lc = L.lcoef + (rowbase + row) * a.NQ;
    int* lo = L.loff + (rowbase + row) * a.NQ;
    int k = 0;
    for (int q = 0; q < a.NQ; q++) {
      const double2 c = oq[q * d + row];
      if (c.x != 0.0 || c.y != 0.0) {
        lc[k] = c;
        lo[k] = a.q_delta[q] * d;
        k++;
      }
    }
    L.lcnt[rowbase + row] = k;
    int* gl = L.glist + (rowbase + row) * a.NG;
    int gk = 0;
    for (int g = 0; g < a.NG; g++) {
      bool any = false;
      for (int p = a.grp_start[g]; p < a.grp_start[g + 1]; p++) {
        const double2 ul = ou[a.grp_pairs[2 * p] * d + row];
        any = any || ul.x != 0.0 || ul.y != 0.0;
      }
      if (any) gl[gk++] = g;
    }
    L.gcnt[rowbase + row] = gk;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// M = A rho + 1/2 sum_j L_j rho L_j^+   (so that k = M + M^+: the two-sided part is Hermitian by itself).
// grid (ceil(d/8), B), 8 warps per block, one output row per warp; NCH = column chunks of 32 per pass.
template <int NCH>
__global__ void __launch_bounds__(256) lindblad_spmm_kernel(const LbArgs a, const LbLists L, const double2* __restrict__ cu, int set,
                                                            const double2* __restrict__ in_all, double2* __restrict__ m_all) {
  __shared__ double2 s_c[8][LB_MAXQ];
  __shared__ int s_o[8][LB_MAXQ];
  __shared__ double2 s_ul[8][LB_MAXQ];  // 1/2 * row factors of the two-sided classes
  __shared__ int s_g[8][LB_MAXQ];       // gather groups acting on the row
  __shared__ int s_gs[LB_MAXQ + 1], s_goff[LB_MAXQ], s_gp[4 * LB_MAXQ];
  const int d = a.d, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long b = blockIdx.y;
  if (threadIdx.x <= a.NG) s_gs[threadIdx.x] = a.grp_start[threadIdx.x];
  if (threadIdx.x < 2 * a.NP) s_gp[threadIdx.x] = a.grp_pairs[threadIdx.x];
  if (threadIdx.x < a.NG) {
    const int p0 = a.grp_start[threadIdx.x];
    s_goff[threadIdx.x] = a.u_delta[a.grp_pairs[2 * p0]] * d + a.u_delta[a.grp_pairs[2 * p0 + 1]];
  }
  __syncthreads();
  // A block walks `rows_per_block` consecutive rows (27 = one block of the three lowest sites for qutrits): most source
  // rows r + delta of its gathers then belong to the same block and are served by L1 instead of L2.
  const int r_end = min(d, (int)(blockIdx.x + 1) * a.rows_per_block);
  for (int r = blockIdx.x * a.rows_per_block + warp; r < r_end; r += 8) {
  __syncwarp();
  const long long rowid = ((long long)set * a.B + b) * d + r;
  const double2* gu = cu + ((long long)set * a.B + b) * a.NU * d;
  const int cnt = L.lcnt[rowid], gc = L.gcnt[rowid];
  if (lane < cnt) {
    s_c[warp][lane] = L.lcoef[rowid * a.NQ + lane];
    s_o[warp][lane] = L.loff[rowid * a.NQ + lane];
  }
  if (lane < gc) s_g[warp][lane] = L.glist[rowid * a.NG + lane];
  if (lane < a.NU) {
    const double2 u = gu[lane * d + r];
    s_ul[warp][lane] = make_double2(0.5 * u.x, 0.5 * u.y);
  }
  __syncwarp();
  const double2* prow = in_all + b * (long long)d * d + (long long)r * d;
  double2* orow = m_all + b * (long long)d * d + (long long)r * d;
  for (int c0 = 0; c0 < d; c0 += 32 * NCH) {
    double ax[NCH], ay[NCH];
#pragma unroll
    for (int ch = 0; ch < NCH; ch++) ax[ch] = ay[ch] = 0.0;
    for (int k = 0; k < cnt; k++) {
      const double2 cf = s_c[warp][k];
      const double2* src = prow + s_o[warp][k] + c0 + lane;
#pragma unroll
      for (int ch = 0; ch < NCH; ch++) {
        if (c0 + ch * 32 + lane < d) {
          const double2 v = __ldg(src + ch * 32);
          ax[ch] = fma(cf.x, v.x, fma(-cf.y, v.y, ax[ch]));
          ay[ch] = fma(cf.x, v.y, fma(cf.y, v.x, ay[ch]));
        }
      }
    }
    if (a.cdense) {
      const double2* cd = a.cdense + (long long)r * d + c0 + lane;
      const double2* src = prow + c0 + lane;
#pragma unroll
      for (int ch = 0; ch < NCH; ch++) {
        if (c0 + ch * 32 + lane < d) {
          const double2 cf = __ldg(cd + ch * 32);
          const double2 v = __ldg(src + ch * 32);
          ax[ch] = fma(cf.x, v.x, fma(-cf.y, v.y, ax[ch]));
          ay[ch] = fma(cf.x, v.y, fma(cf.y, v.x, ay[ch]));
        }
      }
    }
    for (int k = 0; k < gc; k++) {
      const int g = s_g[warp][k];
      const double2* src = prow + s_goff[g] + c0 + lane;
      const int p0 = s_gs[g], p1 = s_gs[g + 1];
      // pair loop outside, chunks unrolled inside: NCH independent loads in flight per instruction group
      double cx[NCH], cy[NCH];
#pragma unroll
      for (int ch = 0; ch < NCH; ch++) cx[ch] = cy[ch] = 0.0;
      for (int p = p0; p < p1; p++) {
        const double2 ul = s_ul[warp][s_gp[2 * p]];
        const double2* urp = gu + s_gp[2 * p + 1] * d + c0 + lane;
#pragma unroll
        for (int ch = 0; ch < NCH; ch++) {
          if (c0 + ch * 32 + lane < d) {
            const double2 ur = __ldg(urp + ch * 32);  // conj(ur) below
            cx[ch] = fma(ul.x, ur.x, fma(ul.y, ur.y, cx[ch]));
            cy[ch] = fma(ul.y, ur.x, fma(-ul.x, ur.y, cy[ch]));
          }
        }
      }
#pragma unroll
      for (int ch = 0; ch < NCH; ch++) {
        // a non-zero coefficient implies the shifted indices are valid
        if (c0 + ch * 32 + lane < d && (cx[ch] != 0.0 || cy[ch] != 0.0)) {
          const double2 v = __ldg(src + ch * 32);
          ax[ch] = fma(cx[ch], v.x, fma(-cy[ch], v.y, ax[ch]));
          ay[ch] = fma(cx[ch], v.y, fma(cy[ch], v.x, ay[ch]));
        }
      }
    }
#pragma unroll
    for (int ch = 0; ch < NCH; ch++)
      if (c0 + ch * 32 + lane < d) orow[c0 + ch * 32 + lane] = make_double2(ax[ch], ay[ch]);
  }
}  // rows of this block
}

// ---------------------------------------------------------------------------------------------------------------
// k = M + M^+ and the RK4 combination.  grid (nt32*nt32, B), block (32, 8): a 32x32 tile, 4 rows per thread,
// lane = column; the partner tile (J,I) is transposed through shared memory.
struct LbStage {
  const double2* in;   // stage input (rho at this stage)
  const double2* lm;   // M of this stage
  const double2* y;    // step start
  const double2* y1;
  const double2* y2;
  double2* out;
  int stage;  // 1..4
  double h;
};

__global__ void __launch_bounds__(256) lindblad_combine_kernel(int d, const LbStage s) {
  __shared__ double2 s_T[32][33];
  const int nt = (d + 31) / 32;
  const int I = blockIdx.x / nt, J = blockIdx.x - I * nt;
  const long long base = (long long)blockIdx.y * d * d;
  const int rI = I * 32, rJ = J * 32;
  const int lane = threadIdx.x, ty = threadIdx.y;
  const double2* lm = s.lm + base;
#pragma unroll
  for (int u = 0; u < 4; u++) {
    const int jl = ty + 8 * u;
    if (rJ + jl < d && rI + lane < d) s_T[jl][lane] = lm[(long long)(rJ + jl) * d + rI + lane];
  }
  __syncthreads();
  const int c = rJ + lane;
  if (c >= d) return;
#pragma unroll
  for (int u = 0; u < 4; u++) {
    const int rl = ty + 8 * u, r = rI + rl;
    if (r >= d) continue;
    const long long o = (long long)r * d + c;
    const double2 l1 = lm[o];
    const double2 lt = s_T[lane][rl];
    const double2 y0 = s.y[base + o];
    const double kx = l1.x + lt.x, ky = l1.y - lt.y;
    double2 res;
    if (s.stage == 1 || s.stage == 2) {
      res = make_double2(fma(0.5 * s.h, kx, y0.x), fma(0.5 * s.h, ky, y0.y));
    } else if (s.stage == 3) {
      res = make_double2(fma(s.h, kx, y0.x), fma(s.h, ky, y0.y));
    } else {
      const double2 v1 = s.y1[base + o], v2 = s.y2[base + o], v3 = s.in[base + o];
      const double t3 = 1.0 / 3.0, h6 = s.h * (1.0 / 6.0);
      res = make_double2(fma(h6, kx, t3 * (v1.x + 2.0 * v2.x + v3.x - y0.x)), fma(h6, ky, t3 * (v1.y + 2.0 * v2.y + v3.y - y0.y)));
    }
    s.out[base + o] = res;
  }
}

// rho0 broadcast / copy into the working buffer
__global__ void lindblad_init_kernel(long long n_per, long long B, const double2* __restrict__ rho0, int batched,
                                     double2* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_per * B) return;
  y[i] = batched ? rho0[i] : rho0[i % n_per];
}

// Observables at one output time: dressed-basis populations p_i = (V^+ rho_lab V)_ii with
// rho_lab[a,b] = e^{-i(e_a-e_b)t} rho_F[a,b].  One block per (state i, trajectory b).
__global__ void __launch_bounds__(256) lindblad_pops_kernel(int d, const double2* __restrict__ rho, const double2* __restrict__ V,
                                                            const double2* __restrict__ ph /*[d] e^{-i e t}*/,
                                                            double* __restrict__ pops, int T, int t_slot) {
  const int i = blockIdx.x;
  const long long b = blockIdx.y;
  const double2* R = rho + b * (long long)d * d;
  __shared__ double red[256];
  double acc = 0.0;
  for (int arow = threadIdx.x; arow < d; arow += blockDim.x) {
    double sx = 0.0, sy = 0.0;
    for (int bc = 0; bc < d; bc++) {
      const double2 v = V[bc * d + i], p = ph[bc];
      const double2 w = make_double2(v.x * p.x + v.y * p.y, v.y * p.x - v.x * p.y);  // V[b,i] conj(ph_b)
      const double2 x = R[(long long)arow * d + bc];
      sx += x.x * w.x - x.y * w.y;
      sy += x.x * w.y + x.y * w.x;
    }
    const double2 va = V[arow * d + i], pa = ph[arow];
    const double2 l = make_double2(va.x * pa.x + va.y * pa.y, va.x * pa.y - va.y * pa.x);  // conj(V[a,i]) ph_a
    acc += l.x * sx - l.y * sy;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) pops[(b * T + t_slot) * (long long)d + i] = red[0];
}

__global__ void lindblad_copy_out_kernel(long long n_per, long long B, const double2* __restrict__ y, double2* __restrict__ out,
                                         int T, int t_slot) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_per * B) return;
  const long long b = i / n_per, e = i - b * n_per;
  out[(b * T + t_slot) * n_per + e] = y[i];
}
