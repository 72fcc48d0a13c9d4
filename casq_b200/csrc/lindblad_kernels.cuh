// casq_b200: device side of the structured Lindblad RK4 path (see lindblad.cu for the host side and the maths).
//
// Per RK4 stage:   k = Lm + Lm^+ + sum_j L_j rho L_j^+      with   Lm = A(t) rho,   A = -i H_F(t) - 1/2 D_F(t)
// where A is a SPARSE d x d matrix (local operators on a product basis: ~6 non-zeros per row for ibmq_manila)
// whose elements are value x time-function look-ups ("shift classes", lindblad.cu).
//
//   lindblad_coef_kernel     once per step: per trajectory and stage time, the class coefficient vectors and, per
//                            row, the COMPACTED list of (coefficient, source-row offset) of the classes acting on it.
//   lindblad_stage_kernel    (lindblad_stage.cuh) one launch per RK4 stage: k = M + M^+ on 64x64 tile pairs, fused
//                            with the RK4 combination.
//
// Round-1 history (profiles/r01_ncu_lindblad_*.txt): (1) tile-pair kernels that gathered per 27x27 tile spent >90 %
// of their instructions on per-tile set-up and index arithmetic and ran at 75-85 traj/s whatever the block shape;
// (2) a row-wise SpMM (M = A rho, one warp per row) followed by a transposing combine kernel reached 145 traj/s with
// 27 matrix passes through HBM per step; (3) the fused stage kernel replaced both (156 traj/s, ~11 passes).
#pragma once
#include "common.cuh"
#include "envelope.cuh"

#define LB_MAXK 4
#define LB_MAXQ 32  // one-sided classes (list slots per row)

struct LbArgs {
  int d;
  long long Bp, b0;  // stride of the parameter block (the whole batch) and first trajectory of this part of it
  int ld;  // row stride of the density-matrix buffers: d rounded up to 8 elements, so every row starts on a 128 B line
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
  const int* pair_ij;      // [NT(NT+1)/2][2] tile pairs (I <= J) of the fused stage kernel
  const double2* cdense;  // [d*d] constant coefficient of the two-sided terms that gather rho[r,c] itself (x 1/2), or NULL
};

struct LbLists {   // per (set, trajectory, row), written by lindblad_coef_kernel
  int* lcnt;       // [3][B][d]       number of one-sided classes acting on the row
  double2* lcoef;  // [3][B][d][NQ]   their coefficients
  int* loff;       // [3][B][d][NQ]   element offset delta*ld of the source row
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
        lo[k] = a.q_delta[q] * a.ld;
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

// One RK4 stage of the fused kernel (lindblad_stage.cuh): four rolling buffers y, y1, y2, y3 and no accumulator:
//   y1 = y + h/2 k(y), y2 = y + h/2 k(y1), y3 = y + h k(y2), y <- (y1 + 2 y2 + y3 - y)/3 + h/6 k(y3).
struct LbStage {
  const double2* in;   // stage input (rho at this stage): the gather source
  const double2* y;    // step start
  const double2* y1;
  const double2* y2;
  double2* out;
  int stage;  // 1..4
  double h;
};

// rho0 broadcast / copy into the (row-padded) working buffer
// Element (r, c) of a density matrix sits at lb_idx(r, c) inside its [d * ld] buffer.  The generic path stores rows
// contiguously (tw = 1: r * ld + c); the chain path stores TILE-MAJOR (tw = 27, tpad = LC_TS - 27): 27 x 27 tiles, each
// stored as 27 rows of tw + tpad elements and contiguous in memory, tiles in row-major order -- a nine-row group of a tile
// is one contiguous piece (3.9 kB), which one bulk asynchronous copy moves.
__device__ __forceinline__ long long lb_idx(int r, int c, int ld, int tw, int tpad) {
  if (tw <= 1) return (long long)r * ld + c;
  const int ts = tw + tpad, nt = ld / ts;
  return ((long long)(r / tw) * nt + c / tw) * (tw * ts) + (r % tw) * ts + c % tw;
}

__global__ void lindblad_init_kernel(int d, int ld, long long B, const double2* __restrict__ rho0, int batched,
                                     double2* __restrict__ y, int tw = 1, int tpad = 0) {
  const long long n_per = (long long)d * d;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_per * B) return;
  const long long b = i / n_per, e = i - b * n_per;
  const int r = (int)(e / d), c = (int)(e - (long long)r * d);
  y[b * (long long)d * ld + lb_idx(r, c, ld, tw, tpad)] = batched ? rho0[i] : rho0[e];
}

// Observables at one output time: dressed-basis populations p_i = (V^+ rho_lab V)_ii with
// rho_lab[a,b] = e^{-i(e_a-e_b)t} rho_F[a,b].  One block per (state i, trajectory b).
__global__ void __launch_bounds__(256) lindblad_pops_kernel(int d, int ld, const double2* __restrict__ rho, const double2* __restrict__ V,
                                                            const double2* __restrict__ ph /*[d] e^{-i e t}*/,
                                                            double* __restrict__ pops, int T, int t_slot, int tw = 1, int tpad = 0) {
  const int i = blockIdx.x;
  const long long b = blockIdx.y;
  const double2* R = rho + b * (long long)d * ld;
  __shared__ double red[256];
  double acc = 0.0;
  for (int arow = threadIdx.x; arow < d; arow += blockDim.x) {
    double sx = 0.0, sy = 0.0;
    for (int bc = 0; bc < d; bc++) {
      const double2 v = V[bc * d + i], p = ph[bc];
      const double2 w = make_double2(v.x * p.x + v.y * p.y, v.y * p.x - v.x * p.y);  // V[b,i] conj(ph_b)
      const double2 x = R[lb_idx(arow, bc, ld, tw, tpad)];
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

__global__ void lindblad_copy_out_kernel(int d, int ld, long long B, const double2* __restrict__ y, double2* __restrict__ out,
                                         int T, int t_slot, int tw = 1, int tpad = 0) {
  const long long n_per = (long long)d * d;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_per * B) return;
  const long long b = i / n_per, e = i - b * n_per;
  const int r = (int)(e / d), c = (int)(e - (long long)r * d);
  out[(b * T + t_slot) * n_per + e] = y[b * (long long)d * ld + lb_idx(r, c, ld, tw, tpad)];
}
