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
// Per RK4 stage ONE kernel sweeps all B density matrices: a tile of rho per block, per-tile coefficient vectors in
// shared memory (so the inner loop is one LDS + one gathered LDG + one complex FMA per class), RK4 combination
// fused into the epilogue (4 rolling buffers, 13 element transfers per step instead of 16).
#include <algorithm>
#include <cstring>
#include <map>

#include "common.cuh"
#include "envelope.cuh"

#define LB_MAXK 4
#define LB_THREADS 256

#define LB_ELEMS 3  // tile elements per thread held in registers: TS*TS <= LB_ELEMS*LB_THREADS

struct LbArgs {
  int d, TS, nt;
  long long B;
  int KA, NF, NTF, NQ, NU, NG, NP, W;
  // tables (device)
  const int* pair_I;      // [n_pairs] tile row index  (I <= J)
  const int* pair_J;      // [n_pairs]
  const int* q_delta;     // [NQ]
  const short* q_code;    // [NQ][d]  time-function id or -1
  const double2* q_val;   // [NQ][d]
  const int* u_delta;     // [NU]
  const short* u_code;    // [NU][d]
  const double2* u_val;   // [NU][d]
  const int* grp_start;   // [NG+1] two-sided pairs grouped by (delta1, delta2): they share one gathered element
  const int* grp_pairs;   // [NP][2]
  const int* tf_sig;      // [NTF] 0 const, 1+k: z_k, 1+KA+k: conj z_k, 1+2KA+k: Re z_k
  const int* tf_freq;     // [NTF] index into the phase part of the stage table, -1 = no phase
  const double* table;    // [NT][W]: KA carriers (cos,sin) then NF phases (cos,sin)
  const int* idx;         // [NT] sample index
  long long entry;        // stage-table entry of this launch
  const double* params;
  DevSlots slots;
  // RK4 buffers
  const double2* in;   // stage input (gathered)
  const double2* y;    // step start
  const double2* y1;
  const double2* y2;
  double2* out;
  int stage;  // 1..4
  double h;
};

__device__ __forceinline__ double2 lb_cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// Per-step coefficient vectors, computed ONCE per (trajectory, stage time) instead of once per tile block:
//   cq[set][b][q][r] = value_q(r) * timefunc_q(r)(t_set)   (one-sided classes, 0 where the class does not act on row r)
//   cu[set][b][q][r] likewise for the two-sided classes.    set 0/1/2 = t, t + h/2, t + h of this step.
struct LbLists {        // per (set, trajectory, row) compacted class lists, written by lindblad_coef_kernel
  int* lcnt;            // [3][B][d]       number of one-sided classes acting on the row
  double2* lcoef;       // [3][B][d][NQ]   their coefficients
  int* loff;            // [3][B][d][NQ]   element offset delta*d of the gathered row
  int* gcnt;            // [3][B][d]       number of two-sided gather groups with a non-zero row factor
  int* glist;           // [3][B][d][NG]
};

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

// rho is Hermitian, so with Lm[r,c] = sum_b alpha(r,b) rho[b,c] (left gathers only) the one-sided part is
// Lm + Lm^+, and rho' itself is Hermitian: one block owns the tile pair (I,J),(J,I), I <= J.  The kernel is
// LATENCY bound (each gather is an L2/DRAM round trip and a row needs 2-3 dependent batches of them), so a block
// is 32 warps wide and every warp owns at most two tile rows (lane = column, TS <= 32): warps 0-15 compute Lm on
// (I,J), warps 16-31 Lm on (J,I) (stored conj-transposed), then each warp finishes one row of (I,J): two-sided
// terms, RK4 combination, write (I,J) and its conjugate mirror.  Everything that depends on the row -- which
// classes act on it, their coefficients, the gathered row -- is warp-uniform and comes pre-compacted from
// lindblad_coef_kernel.  (Round-1 history: v1 spent 90 % of its instructions on index arithmetic and zero tests;
// v3 fixed that but still ran 4 rows x 3 phases serially per warp: 30 us per block whatever the batch size.)
#define LB_STAGE_THREADS 1024

__device__ __forceinline__ void lb_row_gathers(int cnt, const double2* __restrict__ lc, const int* __restrict__ lo,
                                               const double2* __restrict__ pin, double& kx, double& ky) {
  int k = 0;
  for (; k + 4 <= cnt; k += 4) {
    const double2 c0 = lc[k], c1 = lc[k + 1], c2 = lc[k + 2], c3 = lc[k + 3];
    const double2 v0 = __ldg(pin + lo[k]), v1 = __ldg(pin + lo[k + 1]), v2 = __ldg(pin + lo[k + 2]), v3 = __ldg(pin + lo[k + 3]);
    kx = fma(c0.x, v0.x, fma(-c0.y, v0.y, kx));
    ky = fma(c0.x, v0.y, fma(c0.y, v0.x, ky));
    kx = fma(c1.x, v1.x, fma(-c1.y, v1.y, kx));
    ky = fma(c1.x, v1.y, fma(c1.y, v1.x, ky));
    kx = fma(c2.x, v2.x, fma(-c2.y, v2.y, kx));
    ky = fma(c2.x, v2.y, fma(c2.y, v2.x, ky));
    kx = fma(c3.x, v3.x, fma(-c3.y, v3.y, kx));
    ky = fma(c3.x, v3.y, fma(c3.y, v3.x, ky));
  }
  for (; k < cnt; k++) {
    const double2 c0 = lc[k];
    const double2 v0 = __ldg(pin + lo[k]);
    kx = fma(c0.x, v0.x, fma(-c0.y, v0.y, kx));
    ky = fma(c0.x, v0.y, fma(c0.y, v0.x, ky));
  }
}

__global__ void __launch_bounds__(LB_STAGE_THREADS, 1) lindblad_stage_kernel(const LbArgs a, const double2* __restrict__ cu,
                                                                            const LbLists L, int set) {
  const int d = a.d, TS = a.TS, NQ = a.NQ, NU = a.NU, NG = a.NG;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* s_lcA = reinterpret_cast<double2*>(smem_raw);  // [TS][NQ] compacted coefficients, rows of tile I
  double2* s_lcB = s_lcA + TS * NQ;                       // [TS][NQ] rows of tile J
  double2* s_uL = s_lcB + TS * NQ;                        // [NU][TS] two-sided, rows of tile I
  double2* s_uR = s_uL + NU * TS;                         // [NU][TS] two-sided, cols of tile J (conjugated)
  double2* s_L1 = s_uR + NU * TS;                         // [TS][TS+1] Lm of tile (I,J)
  double2* s_LT = s_L1 + TS * (TS + 1);                   // [TS][TS+1] conj-transposed Lm of tile (J,I)
  int* s_loA = reinterpret_cast<int*>(s_LT + TS * (TS + 1));  // [TS][NQ] element offsets delta*d
  int* s_loB = s_loA + TS * NQ;
  int* s_cntA = s_loB + TS * NQ;  // [TS]
  int* s_cntB = s_cntA + TS;      // [TS]
  int* s_gl = s_cntB + TS;        // [TS][NG] groups whose row factor is non-zero
  int* s_gcnt = s_gl + TS * NG;   // [TS]
  int* s_goff = s_gcnt + TS;      // [NG] element offset delta1*d + delta2 of a group
  int* s_gs = s_goff + NG;        // [NG+1] group -> pair range
  int* s_gp = s_gs + NG + 1;      // [2*NP] class pairs

  const long long b = blockIdx.y;
  const int I = a.pair_I[blockIdx.x], J = a.pair_J[blockIdx.x];
  const int rI = I * TS, rJ = J * TS;
  const int nI = min(TS, d - rI), nJ = min(TS, d - rJ);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthreads = blockDim.x, nwarps = nthreads >> 5, half = nwarps >> 1;

  const long long rowbase = ((long long)set * a.B + b) * d;
  const double2* gu = cu + ((long long)set * a.B + b) * NU * d;
  // setup without integer divisions: warp w fetches the lists of row w of I and of J (lane = list slot)
  for (int l = warp; l < TS; l += nwarps) {
    const int rowA = rI + l, rowB = rJ + l;
    const int cA = rowA < d ? L.lcnt[rowbase + rowA] : 0, cB = rowB < d ? L.lcnt[rowbase + rowB] : 0;
    const int gA = rowA < d ? L.gcnt[rowbase + rowA] : 0;
    if (lane < cA) {
      s_lcA[l * NQ + lane] = L.lcoef[(rowbase + rowA) * NQ + lane];
      s_loA[l * NQ + lane] = L.loff[(rowbase + rowA) * NQ + lane];
    }
    if (lane < cB) {
      s_lcB[l * NQ + lane] = L.lcoef[(rowbase + rowB) * NQ + lane];
      s_loB[l * NQ + lane] = L.loff[(rowbase + rowB) * NQ + lane];
    }
    if (lane < gA) s_gl[l * NG + lane] = L.glist[(rowbase + rowA) * NG + lane];
    if (lane == 0) {
      s_cntA[l] = cA;
      s_cntB[l] = cB;
      s_gcnt[l] = gA;
    }
  }
  for (int q = warp; q < NU; q += nwarps)
    if (lane < TS) {
      s_uL[q * TS + lane] = (rI + lane < d) ? gu[q * d + rI + lane] : make_double2(0.0, 0.0);
      const double2 w = (rJ + lane < d) ? gu[q * d + rJ + lane] : make_double2(0.0, 0.0);
      s_uR[q * TS + lane] = make_double2(w.x, -w.y);
    }
  if (tid <= NG) s_gs[tid] = a.grp_start[tid];
  if (tid < 2 * a.NP) s_gp[tid] = a.grp_pairs[tid];
  if (tid < NG) {
    const int p0 = a.grp_start[tid];
    s_goff[tid] = a.u_delta[a.grp_pairs[2 * p0]] * d + a.u_delta[a.grp_pairs[2 * p0 + 1]];
  }
  __syncthreads();

  const double2* in = a.in + b * (long long)d * d;
  // phase A: first half of the warps: Lm on (I,J); second half: Lm on (J,I), stored conj-transposed
  if (warp < half) {
    for (int rl = warp; rl < nI; rl += half)
      if (lane < nJ) {
        double kx = 0.0, ky = 0.0;
        lb_row_gathers(s_cntA[rl], s_lcA + rl * NQ, s_loA + rl * NQ, in + (long long)(rI + rl) * d + rJ + lane, kx, ky);
        s_L1[rl * (TS + 1) + lane] = make_double2(kx, ky);
        if (I == J) s_LT[lane * (TS + 1) + rl] = make_double2(kx, -ky);
      }
  } else if (I != J) {
    for (int rl = warp - half; rl < nJ; rl += half)  // row in J, lane = column in I
      if (lane < nI) {
        double kx = 0.0, ky = 0.0;
        lb_row_gathers(s_cntB[rl], s_lcB + rl * NQ, s_loB + rl * NQ, in + (long long)(rJ + rl) * d + rI + lane, kx, ky);
        s_LT[lane * (TS + 1) + rl] = make_double2(kx, -ky);
      }
  }
  __syncthreads();
  // phase B: k = Lm + Lm^+ + two-sided terms; RK4 combination; write (I,J) and its mirror
  const long long base = b * (long long)d * d;
  // one row per warp (the block has at least TS warps)
  const int rl = warp, cl = lane;
  const bool act = rl < nI && cl < nJ;
  double2 l1 = make_double2(0.0, 0.0), lt = l1;
  if (act) {
    l1 = s_L1[rl * (TS + 1) + cl];
    lt = s_LT[rl * (TS + 1) + cl];
  }
  __syncthreads();  // every (Lm, Lm^+) pair is in registers: s_L1 can take the transposed results
  if (act) {
    const int r = rI + rl, c = rJ + cl;
    const long long o = base + (long long)r * d + c;
    const double2* pin = in + (long long)r * d + c;
    // issue the epilogue loads first: they overlap the two-sided gathers
    const double2 y0 = a.y[o];
    double2 v1 = make_double2(0.0, 0.0), v2 = v1, v3 = v1;
    if (a.stage == 4) {
      v1 = a.y1[o];
      v2 = a.y2[o];
      v3 = pin[0];
    }
    double kx = l1.x + lt.x, ky = l1.y + lt.y;
    const int gc = s_gcnt[rl];
    for (int k0 = 0; k0 < gc; k0 += 4) {
      double cx[4], cy[4];
      double2 v[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        cx[j] = cy[j] = 0.0;
        if (k0 + j < gc) {
          const int g = s_gl[rl * NG + k0 + j];
          for (int p = s_gs[g]; p < s_gs[g + 1]; p++) {
            const double2 ul = s_uL[s_gp[2 * p] * TS + rl], ur = s_uR[s_gp[2 * p + 1] * TS + cl];
            cx[j] = fma(ul.x, ur.x, fma(-ul.y, ur.y, cx[j]));
            cy[j] = fma(ul.x, ur.y, fma(ul.y, ur.x, cy[j]));
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; j++) {  // a non-zero coefficient implies the shifted indices are valid
        v[j] = make_double2(0.0, 0.0);
        if (cx[j] != 0.0 || cy[j] != 0.0) v[j] = __ldg(pin + s_goff[s_gl[rl * NG + k0 + j]]);
      }
#pragma unroll
      for (int j = 0; j < 4; j++) {
        kx = fma(cx[j], v[j].x, fma(-cy[j], v[j].y, kx));
        ky = fma(cx[j], v[j].y, fma(cy[j], v[j].x, ky));
      }
    }
    double2 res;
    if (a.stage == 1 || a.stage == 2) {
      res = make_double2(fma(0.5 * a.h, kx, y0.x), fma(0.5 * a.h, ky, y0.y));
    } else if (a.stage == 3) {
      res = make_double2(fma(a.h, kx, y0.x), fma(a.h, ky, y0.y));
    } else {
      const double t3 = 1.0 / 3.0, h6 = a.h * (1.0 / 6.0);
      res = make_double2(fma(h6, kx, t3 * (v1.x + 2.0 * v2.x + v3.x - y0.x)), fma(h6, ky, t3 * (v1.y + 2.0 * v2.y + v3.y - y0.y)));
    }
    a.out[o] = res;
    if (I != J) s_L1[cl * (TS + 1) + rl] = make_double2(res.x, -res.y);  // own slot was consumed above; see below
  }
  // mirror tile (J,I): written row by row from the transposed copy so the stores are coalesced (27 lanes x 16 B
  // contiguous) instead of 27 scattered partial-sector writes per warp
  if (I != J) {
    __syncthreads();
    for (int rl = warp; rl < nJ; rl += nwarps)
      if (lane < nI) a.out[base + (long long)(rJ + rl) * d + rI + lane] = s_L1[rl * (TS + 1) + lane];
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
  // sum_{a,b'} conj(V[a,i]) ph_a rho[a,b'] conj(ph_b') V[b',i]  (real for Hermitian rho)
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
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
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

__global__ void sv_general_table_kernel(const double* __restrict__ times, long long NT, int KA, int d, int W,
                                        const double* __restrict__ two_pi_nu, const double* __restrict__ fe,
                                        double* __restrict__ out);
int casq_prepare_slots(const casq_model* m, int n_slots, const casq_pulse_slot* slots, DevSlots* ds, int* active,
                       int* n_active, int* n_total);

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

  // ---------------- square tiles; a block owns the tile pair (I,J),(J,I)
  int TS = std::min(d, 27);
  const int NG_ = std::max(1, (int)pairs.size() / 2);  // upper bound on the number of two-sided gather groups
  auto smem_for = [&](int ts) {
    return (size_t)(NQ + NU) * 2 * ts * 16 + 2 * (size_t)ts * (ts + 1) * 16 + ((size_t)2 * ts * NQ + 3 * ts + (size_t)ts * NG_ + 4 * NG_ + 16) * 4 + 64;
  };
  while (smem_for(TS) > 160 * 1024 && TS > 4) TS = (TS + 1) / 2;
  if (smem_for(TS) > 200 * 1024)
    CASQ_FAIL(CASQ_ERR_UNSUPPORTED, "operators too dense for the structured Lindblad kernel (%d one-sided, %d two-sided classes)", NQ, NU);
  const int nt = (d + TS - 1) / TS;
  std::vector<int> pair_I, pair_J;
  for (int I = 0; I < nt; I++)
    for (int Jt = I; Jt < nt; Jt++) {
      pair_I.push_back(I);
      pair_J.push_back(Jt);
    }
  const int n_tile_pairs = (int)pair_I.size();
  // group the two-sided class pairs by their shifts: members of a group gather the same rho element
  std::vector<int> grp_start, grp_pairs;
  {
    std::map<std::pair<int, int>, std::vector<int>> groups;
    for (int p = 0; p < NP; p++) groups[{us.delta[pairs[2 * p]], us.delta[pairs[2 * p + 1]]}].push_back(p);
    grp_start.push_back(0);
    for (auto& kv : groups) {
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
  const size_t o_pI = off; off = align16(off + n_tile_pairs * sizeof(int));
  const size_t o_pJ = off; off = align16(off + n_tile_pairs * sizeof(int));
  const size_t o_tfsig = off; off = align16(off + std::max(NTF, 1) * sizeof(int));
  const size_t o_tffreq = off; off = align16(off + std::max(NTF, 1) * sizeof(int));
  const size_t o_V = off; off = align16(off + (size_t)d * d * 16);
  const size_t o_ph = off; off = align16(off + (size_t)d * 16);
  std::vector<unsigned char> blob(off, 0);
  memcpy(&blob[o_qdelta], os.delta.data(), NQ * sizeof(int));
  memcpy(&blob[o_qcode], os.code.data(), (size_t)NQ * d * sizeof(short));
  memcpy(&blob[o_qval], os.val.data(), (size_t)NQ * d * 16);
  if (NU) {
    memcpy(&blob[o_udelta], us.delta.data(), NU * sizeof(int));
    memcpy(&blob[o_ucode], us.code.data(), (size_t)NU * d * sizeof(short));
    memcpy(&blob[o_uval], us.val.data(), (size_t)NU * d * 16);
  }
  if (NP) memcpy(&blob[o_pairs], grp_pairs.data(), (size_t)NP * 2 * sizeof(int));
  memcpy(&blob[o_gstart], grp_start.data(), (NG + 1) * sizeof(int));
  memcpy(&blob[o_pI], pair_I.data(), n_tile_pairs * sizeof(int));
  memcpy(&blob[o_pJ], pair_J.data(), n_tile_pairs * sizeof(int));
  if (NTF) {
    memcpy(&blob[o_tfsig], tf_sig.data(), NTF * sizeof(int));
    memcpy(&blob[o_tffreq], tf_freq.data(), NTF * sizeof(int));
  }
  memcpy(&blob[o_V], m->V.data(), (size_t)d * d * 16);

  const size_t n_per = (size_t)d * d;
  if ((rc = m->tab_times.ensure(NT * sizeof(double)))) return rc;
  if ((rc = m->tab_idx.ensure(NT * sizeof(int)))) return rc;
  if ((rc = m->tab.ensure((size_t)NT * std::max(W, 2) * sizeof(double)))) return rc;
  if ((rc = m->work0.ensure(consts.size() * sizeof(double)))) return rc;
  if ((rc = m->work1.ensure(blob.size()))) return rc;
  if ((rc = m->work2.ensure(4 * (size_t)B * n_per * 16))) return rc;
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
  double2* bufs[4];
  for (int i = 0; i < 4; i++) bufs[i] = (double2*)m->work2.p + (size_t)i * B * n_per;
  {
    const long long n = (long long)B * n_per;
    lindblad_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((long long)n_per, B, (const double2*)rho0_dev, rho0_batched, bufs[0]);
    CASQ_CUDA(cudaGetLastError());
  }
  LbArgs a;
  memset(&a, 0, sizeof(a));
  a.d = d; a.TS = TS; a.nt = nt; a.B = B;
  a.KA = KA; a.NF = NF; a.NTF = NTF; a.NQ = NQ; a.NU = NU; a.NG = NG; a.NP = NP; a.W = W;
  a.pair_I = (const int*)(dev + o_pI);
  a.pair_J = (const int*)(dev + o_pJ);
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
  a.slots = ds;
  const size_t smem = smem_for(TS);
  CASQ_CUDA(cudaFuncSetAttribute(lindblad_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const size_t smem_coef = (size_t)std::max(NTF, 1) * 16;
  CASQ_CUDA(cudaFuncSetAttribute(lindblad_coef_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_coef));
  const dim3 grid((unsigned)n_tile_pairs, (unsigned)B);
  if (B > 65535) CASQ_FAIL(CASQ_ERR_UNSUPPORTED, "casq_solve_lindblad_rk4: at most 65535 density matrices per call");

  std::vector<cplx> ph(d);
  int t_slot = 0;
  auto emit = [&](double t) -> int {
    if (out_rho_dev) {
      const long long n = (long long)B * n_per;
      lindblad_copy_out_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((long long)n_per, B, bufs[0], (double2*)out_rho_dev, T, t_slot);
      CASQ_CUDA(cudaGetLastError());
    }
    if (out_pops_dev) {
      for (int i = 0; i < d; i++) ph[i] = std::polar(1.0, -m->fe[i] * t);
      // synchronous small copy: ph is reused by the next output point
      CASQ_CUDA(cudaMemcpyAsync(dev + o_ph, ph.data(), (size_t)d * 16, cudaMemcpyHostToDevice, st));
      CASQ_CUDA(cudaStreamSynchronize(st));
      lindblad_pops_kernel<<<dim3((unsigned)d, (unsigned)B), 256, 0, st>>>(d, bufs[0], (const double2*)(dev + o_V), (const double2*)(dev + o_ph),
                                                                        out_pops_dev, T, t_slot);
      CASQ_CUDA(cudaGetLastError());
    }
    t_slot++;
    return CASQ_OK;
  };
  if (g.out_keep[0] && (rc = emit(g.out_times[0]))) return rc;
  long long pos = 0;
  for (size_t p = 0; p < g.piece_nsteps.size(); p++) {
    const int n = g.piece_nsteps[p];
    a.h = g.piece_h[p];
    for (int s = 0; s < n; s++) {
      // stage 1: in = y (bufs[0]) -> y1 (bufs[1]); 2: in = y1 -> y2; 3: in = y2 -> y3; 4: in = y3 -> y (in place)
      lindblad_coef_kernel<<<dim3((unsigned)B, 3), 256, smem_coef, st>>>(a, pos, cq_all, cu_all, lists);
      for (int stage = 1; stage <= 4; stage++) {
        a.stage = stage;
        const int set = stage == 1 ? 0 : (stage == 4 ? 2 : 1);
        a.entry = pos + set;
        a.in = bufs[stage - 1];
        a.y = bufs[0];
        a.y1 = bufs[1];
        a.y2 = bufs[2];
        a.out = stage == 4 ? bufs[0] : bufs[stage];
        lindblad_stage_kernel<<<grid, LB_STAGE_THREADS, smem, st>>>(a, cu_all, lists, set);
      }
      pos += 2;
    }
    CASQ_CUDA(cudaGetLastError());
    pos += 1;
    if (g.out_keep[p + 1] && (rc = emit(g.out_times[p + 1]))) return rc;
  }
  return CASQ_OK;
}
