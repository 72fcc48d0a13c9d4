// casq_b200: piecewise-constant propagators U = prod_n expm(h G_F(t_n + h/2)) for batches of pulse points,
// one warp per trajectory, all matrices in shared memory, for dim <= 16; a block per trajectory (propagator_block.cuh)
// for 16 < dim <= 64.
//
// This is the exponential-midpoint scheme of qiskit-dynamics' fixed-step "scipy_expm"/"jax_expm" solvers
// (SURVEY.md Appendix A.5 last bullet).  casq's ODESolverMethod enum cannot select them
// (src/casq/backends/pulse_backend.py:60-71), so this entry point is additive API for BASELINE cfg 3
// ("2 coupled transmons (dim 9) cross-resonance GaussianSquare sweep via full-unitary propagator batches").
// expm is the scaling-and-squaring Pade approximant of Higham (2005) -- the algorithm behind
// jax.scipy.linalg.expm: order m in {3,5,7,9} chosen from ||A||_1 (above theta_9 the matrix is scaled to theta_9;
// Higham's [13/13] branch needs three more live matrices, see wexpm), [m/m] Pade solved by Gaussian elimination with
// partial pivoting, then s squarings.  d < 16 here, so this is CUDA-core complex arithmetic;
// a split-complex tensor-core GEMM only pays for d >= 16 (north_star) and is not built.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "envelope.cuh"

#define PROP_MAXK 4
#define PROP_WARPS 4   // warps (= trajectories) per block of the warp-per-trajectory kernel
#define PROP_MAXD 64

struct PropArgs {
  const double* table;  // [NT][W]: carriers then p_i at every stage time (midpoints are the odd entries)
  const int* idx;
  const int* piece_n;
  const double* piece_h;
  int n_pieces;
  long long B;
  int d, KA, W, rwa;
  const double* params;
  double* out;  // [B][d][d] c128
  DevSlots slots;
  const double2* S;  // dense frame-basis operators [d*d], [KA][d*d]
  const double2* P;
  const double2* N;
  int identityU;
  const double2* U;  // frame basis
  const double2* env_tab;  // [KA][n_total][B] envelope samples (casq_env_table_kernel)
  int n_total;
  double2* scratch;  // block kernel, dim > 32: [B][6][dim*dim] work matrices in global memory (NULL: shared memory)
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cfma(double2 a, double2 b, double2 c) {
  return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}

// z_k = envelope sample (per-solve table) x carrier phase at the midpoint entry `pm` of the stage table, for every driven
// channel; shared by the warp-per-trajectory and the block-per-trajectory kernels.
__device__ __forceinline__ void prop_signals(const PropArgs& a, long long b, long long pm, double* zr, double* zi) {
  const int gidx = __ldg(a.idx + pm);
  const bool inside = gidx >= 0 && gidx < a.n_total;
  const double* ent = a.table + pm * a.W;
#pragma unroll
  for (int k = 0; k < PROP_MAXK; k++) {
    zr[k] = zi[k] = 0.0;
    if (k < a.KA && inside) {
      const double2 env = __ldg(a.env_tab + ((long long)k * a.n_total + gidx) * a.B + b);
      const double c = __ldg(ent + 2 * k), sn = __ldg(ent + 2 * k + 1);
      zr[k] = env.x * c - env.y * sn;
      zi[k] = env.x * sn + env.y * c;
    }
  }
}

// A = h * (-i) * p_i conj(p_j) (S + sum_k z_k P_k + conj(z_k) N_k), elements strided over NT cooperating threads.  The
// operators are read through L1 (the same few kB for every trajectory of the grid): shared memory is kept for the work
// matrices.
template <int DT>
__device__ __forceinline__ void prop_assemble(const PropArgs& a, int d_, long long pm, const double* zr, const double* zi, double h,
                                              double2* A, int tid, int nthreads) {
  const int d = DT > 0 ? DT : d_, n2 = d * d;
  const double2* ph = reinterpret_cast<const double2*>(a.table + pm * a.W + 2 * a.KA);
  for (int e = tid; e < n2; e += nthreads) {
    const int i = e / d, j = e - i * d;
    double2 m = __ldg(a.S + e);
#pragma unroll
    for (int k = 0; k < PROP_MAXK; k++) {
      if (k < a.KA) {
        const double2 P = __ldg(a.P + (size_t)k * n2 + e);
        if (a.rwa) {
          const double2 N = __ldg(a.N + (size_t)k * n2 + e);
          m.x += zr[k] * (P.x + N.x) - zi[k] * (P.y - N.y);
          m.y += zr[k] * (P.y + N.y) + zi[k] * (P.x - N.x);
        } else {
          m.x = fma(zr[k], P.x, m.x);
          m.y = fma(zr[k], P.y, m.y);
        }
      }
    }
    const double2 pi_ = __ldg(ph + i), pj = __ldg(ph + j);
    const double2 f = make_double2(pi_.x * pj.x + pi_.y * pj.y, pi_.y * pj.x - pi_.x * pj.y);  // p_i conj(p_j)
    const double2 g = cmul(f, m);
    A[e] = make_double2(h * g.y, -h * g.x);  // -i g
  }
}

// C = X * Y (d x d), warp-cooperative; C must not alias X or Y.
template <int DT>
__device__ __forceinline__ void wmatmul(int d_, int lane, double2* C, const double2* X, const double2* Y) {
  const int d = DT > 0 ? DT : d_;
  if (DT == 9) {
    // 27 lanes, lane = (row i, column block of 3): X[i][k] is loaded once for three outputs (36 LDS per 27 complex
    // FMAs instead of 54)
    if (lane < 27) {
      const int i = lane / 3, j0 = 3 * (lane - 3 * i);
      double2 a0 = make_double2(0.0, 0.0), a1 = a0, a2 = a0;
#pragma unroll
      for (int k = 0; k < 9; k++) {
        const double2 x = X[i * 9 + k];
        a0 = cfma(x, Y[k * 9 + j0], a0);
        a1 = cfma(x, Y[k * 9 + j0 + 1], a1);
        a2 = cfma(x, Y[k * 9 + j0 + 2], a2);
      }
      C[i * 9 + j0] = a0;
      C[i * 9 + j0 + 1] = a1;
      C[i * 9 + j0 + 2] = a2;
    }
    __syncwarp();
    return;
  }
  for (int e = lane; e < d * d; e += 32) {
    const int i = e / d, j = e - i * d;
    double2 acc = make_double2(0.0, 0.0);
    for (int k = 0; k < d; k++) acc = cfma(X[i * d + k], Y[k * d + j], acc);
    C[e] = acc;
  }
  __syncwarp();
}

// Solve Q X = Pm in place (X overwrites Pm); Q is destroyed.  Gaussian elimination, partial pivoting.
template <int DT>
__device__ void wsolve(int d_, int lane, double2* Q, double2* Pm) {
  const int d = DT > 0 ? DT : d_;
  for (int k = 0; k < d; k++) {
    // pivot search over rows k..d-1 of column k
    double best = -1.0;
    int brow = k;
    for (int r = k + lane; r < d; r += 32) {
      const double2 v = Q[r * d + k];
      const double mag = v.x * v.x + v.y * v.y;
      if (mag > best) {
        best = mag;
        brow = r;
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int orow = __shfl_xor_sync(0xffffffffu, brow, o);
      if (ob > best || (ob == best && orow < brow)) {
        best = ob;
        brow = orow;
      }
    }
    if (brow != k) {
      for (int j = lane; j < 2 * d; j += 32) {
        double2* M = (j < d) ? Q : Pm;
        const int c = (j < d) ? j : j - d;
        const double2 t = M[k * d + c];
        M[k * d + c] = M[brow * d + c];
        M[brow * d + c] = t;
      }
      __syncwarp();
    }
    const double2 piv = Q[k * d + k];
    const double inv_n = 1.0 / (piv.x * piv.x + piv.y * piv.y);
    const double2 pinv = make_double2(piv.x * inv_n, -piv.y * inv_n);
    __syncwarp();
    // scale pivot row
    for (int j = lane; j < 2 * d; j += 32) {
      double2* M = (j < d) ? Q : Pm;
      const int c = (j < d) ? j : j - d;
      if (j >= d || c >= k) M[k * d + c] = cmul(M[k * d + c], pinv);
    }
    __syncwarp();
    // eliminate column k from every other row (Gauss-Jordan: no back substitution pass)
    for (int e = lane; e < d * 2 * d; e += 32) {
      const int r = e / (2 * d), j = e - r * (2 * d);
      if (r == k) continue;
      double2* M = (j < d) ? Q : Pm;
      const int c = (j < d) ? j : j - d;
      if (j < d && c <= k) continue;  // those entries become 0/unused; the factor column is read below
      const double2 f = Q[r * d + k];
      const double2 p = M[k * d + c];
      M[r * d + c] = make_double2(M[r * d + c].x - (f.x * p.x - f.y * p.y), M[r * d + c].y - (f.x * p.y + f.y * p.x));
    }
    __syncwarp();
  }
}

// d = 9: the augmented matrix [Q | P] (9 x 18) lives in REGISTERS, lane = (row i, block of 6 columns), 27 lanes.
// Rows are never swapped: the pivot of column k is the largest entry among the rows not used as a pivot yet (partial
// pivoting up to a row permutation) and row i is written to row "variable it solved" of the result at the end.  The
// arg-max is ONE redux.sync on a 32-bit key (upper bits of |x|^2, row index in the low 4 bits: candidates that agree to
// 2^-16 are interchangeable as pivots).  Per pivot: 1 REDUX + 4 SHFL + the pivot row through shared memory, against the
// generic version's five shuffle rounds, physical row swaps and 162 element updates strided over 32 lanes.
__device__ __forceinline__ void wsolve9(int lane, double2* Q, double2* Pm) {
  const int i = lane / 3, jb = lane - 3 * i;
  const bool act = lane < 27;
  double2 v[6];
#pragma unroll
  for (int t = 0; t < 6; t++) {
    const int c = 6 * jb + t;
    v[t] = act ? (c < 9 ? Q[i * 9 + c] : Pm[i * 9 + c - 9]) : make_double2(0.0, 0.0);
  }
  __syncwarp();  // Q is scratch from here on
  double2* s_row = Q;     // [18] scaled pivot row
  double2* s_f = Q + 18;  // [9]  elimination factors (column k)
  int myk = -1;           // the variable this lane's row solved, -1 while the row is still a pivot candidate
#pragma unroll
  for (int k = 0; k < 9; k++) {
    const int kb = k / 6, kt = k % 6;
    const double2 x = v[kt];
    unsigned key = 0;
    if (act && jb == kb && myk < 0) key = ((unsigned)__double2hiint(x.x * x.x + x.y * x.y) & 0xFFFFFFF0u) | (unsigned)(15 - i);
    key = __reduce_max_sync(0xffffffffu, key);
    const int pr = 15 - (int)(key & 15u);
    const double px = __shfl_sync(0xffffffffu, x.x, 3 * pr + kb), py = __shfl_sync(0xffffffffu, x.y, 3 * pr + kb);
    const double inv_n = 1.0 / (px * px + py * py);
    const double2 pinv = make_double2(px * inv_n, -py * inv_n);
    if (act) {
      if (i == pr) {
#pragma unroll
        for (int t = 0; t < 6; t++) {
          v[t] = cmul(v[t], pinv);
          s_row[6 * jb + t] = v[t];
        }
        myk = k;
      } else if (jb == kb) {
        s_f[i] = x;
      }
    }
    __syncwarp();
    if (act && i != pr) {
      const double2 f = s_f[i];
#pragma unroll
      for (int t = 0; t < 6; t++) {
        const double2 p = s_row[6 * jb + t];
        v[t].x = fma(-f.x, p.x, fma(f.y, p.y, v[t].x));
        v[t].y = fma(-f.x, p.y, fma(-f.y, p.x, v[t].y));
      }
    }
    __syncwarp();
  }
  if (act) {
#pragma unroll
    for (int t = 0; t < 6; t++) {
      const int c = 6 * jb + t;
      if (c >= 9) Pm[myk * 9 + c - 9] = v[t];
    }
  }
  __syncwarp();
}

// wsolve9 with the operands already in registers in the PRODUCT layout: lane (row i, column block tj) holds
// Q[i][3tj..3tj+2] and P[i][3tj..3tj+2], i.e. columns {3tj..} and {9+3tj..} of the augmented matrix, so Q = V - U and
// P = V + U go from the product epilogue into the elimination without a round trip through shared memory.
// scratch: 27 double2 (pivot row [18] + factors [9]); the solution is written to Eout (9x9).
__device__ __forceinline__ void wsolve9_tiles(int lane, double2 (&q)[3], double2 (&p)[3], double2* scratch, double2* Eout) {
  const int i = lane / 3, tj = lane - 3 * i;
  const bool act = lane < 27;
  double2* s_row = scratch;
  double2* s_f = scratch + 18;
  int myk = -1;
#pragma unroll
  for (int k = 0; k < 9; k++) {
    const int kb = k / 3, kt = k % 3;
    const double2 x = q[kt];
    unsigned key = 0;
    if (act && tj == kb && myk < 0) key = ((unsigned)__double2hiint(x.x * x.x + x.y * x.y) & 0xFFFFFFF0u) | (unsigned)(15 - i);
    key = __reduce_max_sync(0xffffffffu, key);
    const int pr = 15 - (int)(key & 15u);
    const double px = __shfl_sync(0xffffffffu, x.x, 3 * pr + kb), py = __shfl_sync(0xffffffffu, x.y, 3 * pr + kb);
    const double inv_n = 1.0 / (px * px + py * py);
    const double2 pinv = make_double2(px * inv_n, -py * inv_n);
    if (act) {
      if (i == pr) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
          q[c] = cmul(q[c], pinv);
          p[c] = cmul(p[c], pinv);
          s_row[3 * tj + c] = q[c];
          s_row[9 + 3 * tj + c] = p[c];
        }
        myk = k;
      } else if (tj == kb) {
        s_f[i] = x;
      }
    }
    __syncwarp();
    if (act && i != pr) {
      const double2 f = s_f[i];
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const double2 rq = s_row[3 * tj + c], rp = s_row[9 + 3 * tj + c];
        q[c].x = fma(-f.x, rq.x, fma(f.y, rq.y, q[c].x));
        q[c].y = fma(-f.x, rq.y, fma(-f.y, rq.x, q[c].y));
        p[c].x = fma(-f.x, rp.x, fma(f.y, rp.y, p[c].x));
        p[c].y = fma(-f.x, rp.y, fma(-f.y, rp.x, p[c].y));
      }
    }
    __syncwarp();
  }
  if (act) {
#pragma unroll
    for (int c = 0; c < 3; c++) Eout[myk * 9 + 3 * tj + c] = p[c];
  }
  __syncwarp();
}

// d = 9 product with the output tile (row i, columns j0..j0+2) left in REGISTERS; lanes >= 27 idle.
__device__ __forceinline__ void mm9_tile(int lane, const double2* X, const double2* Y, double2 (&t)[3]) {
  t[0] = t[1] = t[2] = make_double2(0.0, 0.0);
  if (lane < 27) {
    const int i = lane / 3, j0 = 3 * (lane - 3 * i);
#pragma unroll
    for (int k = 0; k < 9; k++) {
      const double2 x = X[i * 9 + k];
      t[0] = cfma(x, Y[k * 9 + j0], t[0]);
      t[1] = cfma(x, Y[k * 9 + j0 + 1], t[1]);
      t[2] = cfma(x, Y[k * 9 + j0 + 2], t[2]);
    }
  }
}
__device__ __forceinline__ void st9_tile(int lane, double2* C, const double2 (&t)[3]) {
  if (lane < 27) {
    const int i = lane / 3, j0 = 3 * (lane - 3 * i);
    C[i * 9 + j0] = t[0];
    C[i * 9 + j0 + 1] = t[1];
    C[i * 9 + j0 + 2] = t[2];
  }
}

// d = 9 version of wexpm with the element-wise passes fused into the product epilogues: the Pade polynomials are
// accumulated from the power tiles while they are still in registers (A^4 / A^8 are never stored, V never leaves the
// registers), and Q = V - U, P = V + U are formed from the U tile.  The kernel is bound by the shared-memory data pipe
// (L1TEX 90 %): per step this removes 9 STS.128 and 12 LDS.128 per lane.
__device__ double2* wexpm9(int lane, double2* A, double2* X1, double2* X2, double2* X3, double2* X4) {
  for (int e = lane; e < 81; e += 32) X4[e].x = sqrt(A[e].x * A[e].x + A[e].y * A[e].y);
  __syncwarp();
  double nrm = 0.0;
  if (lane < 9) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 9; i++) s += X4[i * 9 + lane].x;
    nrm = s;
  }
  for (int o = 16; o > 0; o >>= 1) nrm = fmax(nrm, __shfl_xor_sync(0xffffffffu, nrm, o));
  __syncwarp();
  int mi, s = 0;  // Pade order m = 3 + 2 mi
  if (nrm < 1.495585217958292e-2)
    mi = 0;
  else if (nrm < 2.539398330063230e-1)
    mi = 1;
  else if (nrm < 9.504178996162932e-1)
    mi = 2;
  else {
    mi = 3;
    const double r = nrm / 2.097847961257068e0;
    s = r > 1.0 ? (int)ceil(log2(r)) : 0;
    if (s > 0) {
      const double sc = ldexp(1.0, -s);
      for (int e = lane; e < 81; e += 32) A[e] = make_double2(A[e].x * sc, A[e].y * sc);
      __syncwarp();
    }
  }
  const double B3[4] = {120., 60., 12., 1.};
  const double B5[6] = {30240., 15120., 3360., 420., 30., 1.};
  const double B7[8] = {17297280., 8648640., 1995840., 277200., 25200., 1512., 56., 1.};
  const double B9[10] = {17643225600., 8821612800., 2075673600., 302702400., 30270240., 2162160., 110880., 3960., 90., 1.};
  auto bc = [&](int k) { return mi == 0 ? B3[k] : (mi == 1 ? B5[k] : (mi == 2 ? B7[k] : B9[k])); };
  const int i = lane / 3, j0 = 3 * (lane - 3 * i);
  double2 t[3], od[3], vm[3];
  mm9_tile(lane, A, A, t);  // A^2
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const double dg = (i == j0 + c) ? 1.0 : 0.0;
    od[c] = make_double2(fma(bc(3), t[c].x, bc(1) * dg), bc(3) * t[c].y);
    vm[c] = make_double2(fma(bc(2), t[c].x, bc(0) * dg), bc(2) * t[c].y);
  }
  if (mi >= 1) {
    st9_tile(lane, X1, t);
    __syncwarp();
    mm9_tile(lane, X1, X1, t);  // A^4
#pragma unroll
    for (int c = 0; c < 3; c++) {
      od[c] = make_double2(fma(bc(5), t[c].x, od[c].x), fma(bc(5), t[c].y, od[c].y));
      vm[c] = make_double2(fma(bc(4), t[c].x, vm[c].x), fma(bc(4), t[c].y, vm[c].y));
    }
    if (mi >= 2) {
      st9_tile(lane, X2, t);
      __syncwarp();
      mm9_tile(lane, X2, X1, t);  // A^6
#pragma unroll
      for (int c = 0; c < 3; c++) {
        od[c] = make_double2(fma(bc(7), t[c].x, od[c].x), fma(bc(7), t[c].y, od[c].y));
        vm[c] = make_double2(fma(bc(6), t[c].x, vm[c].x), fma(bc(6), t[c].y, vm[c].y));
      }
      if (mi >= 3) {
        st9_tile(lane, X3, t);
        __syncwarp();
        mm9_tile(lane, X3, X1, t);  // A^8
#pragma unroll
        for (int c = 0; c < 3; c++) {
          od[c] = make_double2(fma(bc(9), t[c].x, od[c].x), fma(bc(9), t[c].y, od[c].y));
          vm[c] = make_double2(fma(bc(8), t[c].x, vm[c].x), fma(bc(8), t[c].y, vm[c].y));
        }
      }
    }
  }
  st9_tile(lane, X4, od);  // the |A| scratch is free; X4 holds the odd polynomial
  __syncwarp();            // also: every lane is done with X1 / X2 / X3 as operands
  mm9_tile(lane, A, X4, t);  // U
  double2 qq[3], pp[3];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    qq[c] = make_double2(vm[c].x - t[c].x, vm[c].y - t[c].y);  // Q = V - U
    pp[c] = make_double2(vm[c].x + t[c].x, vm[c].y + t[c].y);  // P = V + U
  }
  wsolve9_tiles(lane, qq, pp, X3, X1);  // X3: scratch (A^6 is dead); solution -> X1 (A^2 is dead)
  double2 *E = X1, *Sq = X2;
  for (int q = 0; q < s; q++) {
    wmatmul<9>(9, lane, Sq, E, E);
    double2* tmp = E;
    E = Sq;
    Sq = tmp;
  }
  return E;
}

// expm(A) with FOUR work matrices X1..X4 (d*d each); A is overwritten (scaled); returns the buffer holding the result
// (X1 or X2).  Orders 3/5/7/9 by Higham's theta_m; above theta_9 the matrix is scaled down to theta_9 and squared back.
// (Higham switches to [13/13] with theta_13 there: same backward-error bound, but it needs three more live matrices.
// With A, the running product and X1..X4 a warp needs 6 matrices of shared memory instead of 9, which is what lets all
// 4 096 warps of BASELINE cfg 3 be resident at once.)  Buffer life: X1 = A^2 -> U -> P -> result; X2 = A^4 -> squaring
// scratch; X3 = A^6 -> V -> Q; X4 = |A| scratch -> A^8 -> odd polynomial.
template <int DT>
__device__ double2* wexpm(int d_, int lane, double2* A, double2* X1, double2* X2, double2* X3, double2* X4) {
  const int d = DT > 0 ? DT : d_;
  const int n2 = d * d;
  // ||A||_1 = max column sum: element magnitudes spread over the warp (X4 as scratch), then one lane per column
  for (int e = lane; e < n2; e += 32) X4[e].x = sqrt(A[e].x * A[e].x + A[e].y * A[e].y);
  __syncwarp();
  double nrm = 0.0;
  for (int j = lane; j < d; j += 32) {
    double s = 0.0;
    for (int i = 0; i < d; i++) s += X4[i * d + j].x;
    nrm = fmax(nrm, s);
  }
  for (int o = 16; o > 0; o >>= 1) nrm = fmax(nrm, __shfl_xor_sync(0xffffffffu, nrm, o));
  __syncwarp();
  int m, s = 0;
  if (nrm < 1.495585217958292e-2)
    m = 3;
  else if (nrm < 2.539398330063230e-1)
    m = 5;
  else if (nrm < 9.504178996162932e-1)
    m = 7;
  else {
    m = 9;
    const double r = nrm / 2.097847961257068e0;
    s = r > 1.0 ? (int)ceil(log2(r)) : 0;
    if (s > 0) {
      const double sc = ldexp(1.0, -s);
      for (int e = lane; e < n2; e += 32) A[e] = make_double2(A[e].x * sc, A[e].y * sc);
      __syncwarp();
    }
  }
  double2 *A2 = X1, *A4 = X2, *A6 = X3, *A8 = X4, *Od = X4, *Vm = X3;
  wmatmul<DT>(d, lane, A2, A, A);
  auto diag = [&](int e) { return (e / d) == (e % d) ? 1.0 : 0.0; };
  if (m == 3) {
    const double b[4] = {120., 60., 12., 1.};
    for (int e = lane; e < n2; e += 32) {
      Od[e] = make_double2(b[3] * A2[e].x + b[1] * diag(e), b[3] * A2[e].y);
      Vm[e] = make_double2(b[2] * A2[e].x + b[0] * diag(e), b[2] * A2[e].y);
    }
  } else if (m == 5) {
    const double b[6] = {30240., 15120., 3360., 420., 30., 1.};
    wmatmul<DT>(d, lane, A4, A2, A2);
    for (int e = lane; e < n2; e += 32) {
      Od[e] = make_double2(b[5] * A4[e].x + b[3] * A2[e].x + b[1] * diag(e), b[5] * A4[e].y + b[3] * A2[e].y);
      Vm[e] = make_double2(b[4] * A4[e].x + b[2] * A2[e].x + b[0] * diag(e), b[4] * A4[e].y + b[2] * A2[e].y);
    }
  } else if (m == 7) {
    const double b[8] = {17297280., 8648640., 1995840., 277200., 25200., 1512., 56., 1.};
    wmatmul<DT>(d, lane, A4, A2, A2);
    wmatmul<DT>(d, lane, A6, A4, A2);
    for (int e = lane; e < n2; e += 32) {  // V overwrites A^6 element by element
      const double2 a6 = A6[e];
      Od[e] = make_double2(b[7] * a6.x + b[5] * A4[e].x + b[3] * A2[e].x + b[1] * diag(e), b[7] * a6.y + b[5] * A4[e].y + b[3] * A2[e].y);
      Vm[e] = make_double2(b[6] * a6.x + b[4] * A4[e].x + b[2] * A2[e].x + b[0] * diag(e), b[6] * a6.y + b[4] * A4[e].y + b[2] * A2[e].y);
    }
  } else {
    const double b[10] = {17643225600., 8821612800., 2075673600., 302702400., 30270240., 2162160., 110880., 3960., 90., 1.};
    wmatmul<DT>(d, lane, A4, A2, A2);
    wmatmul<DT>(d, lane, A6, A4, A2);
    wmatmul<DT>(d, lane, A8, A6, A2);
    for (int e = lane; e < n2; e += 32) {  // the polynomials overwrite A^8 and A^6 element by element
      const double2 a8 = A8[e], a6 = A6[e];
      Od[e] = make_double2(b[9] * a8.x + b[7] * a6.x + b[5] * A4[e].x + b[3] * A2[e].x + b[1] * diag(e),
                           b[9] * a8.y + b[7] * a6.y + b[5] * A4[e].y + b[3] * A2[e].y);
      Vm[e] = make_double2(b[8] * a8.x + b[6] * a6.x + b[4] * A4[e].x + b[2] * A2[e].x + b[0] * diag(e),
                           b[8] * a8.y + b[6] * a6.y + b[4] * A4[e].y + b[2] * A2[e].y);
    }
  }
  __syncwarp();
  double2* Um = X1;  // A^2 is dead
  wmatmul<DT>(d, lane, Um, A, Od);  // U = A * (odd polynomial part)
  // Q = V - U (over V), P = V + U (over U); solve Q R = P
  for (int e = lane; e < n2; e += 32) {
    const double2 v = Vm[e], u = Um[e];
    X3[e] = make_double2(v.x - u.x, v.y - u.y);
    X1[e] = make_double2(v.x + u.x, v.y + u.y);
  }
  __syncwarp();
  if (DT == 9)
    wsolve9(lane, X3, X1);
  else
    wsolve<DT>(d, lane, X3, X1);
  double2 *E = X1, *Sq = X2;
  for (int q = 0; q < s; q++) {
    wmatmul<DT>(d, lane, Sq, E, E);
    double2* t = E;
    E = Sq;
    Sq = t;
  }
  return E;
}

#define PROP_NBUF 6  // A, running product, X1..X4

template <int DT>
__global__ void __launch_bounds__(32 * PROP_WARPS, DT == 9 ? 7 : 1) propagator_expm_kernel(const PropArgs a) {
  const int d = DT > 0 ? DT : a.d, n2 = d * d;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (b >= a.B) return;
  double2* ws = reinterpret_cast<double2*>(smem_raw) + (size_t)warp * PROP_NBUF * n2;
  double2 *A = ws, *Uacc = ws + n2, *X1 = ws + 2 * n2, *X2 = ws + 3 * n2, *X3 = ws + 4 * n2, *X4 = ws + 5 * n2;
  for (int e = lane; e < n2; e += 32) Uacc[e] = make_double2((e / d) == (e % d) ? 1.0 : 0.0, 0.0);
  __syncwarp();

  long long pos = 0;
  for (int p = 0; p < a.n_pieces; p++) {
    const int n = a.piece_n[p];
    const double h = a.piece_h[p];
    for (int s = 0; s < n; s++) {
      const long long pm = pos + 1;  // midpoint entry
      double zr[PROP_MAXK], zi[PROP_MAXK];
      prop_signals(a, b, pm, zr, zi);
      prop_assemble<DT>(a, d, pm, zr, zi, h, A, lane, 32);
      __syncwarp();
      double2* E = DT == 9 ? wexpm9(lane, A, X1, X2, X3, X4) : wexpm<DT>(d, lane, A, X1, X2, X3, X4);
      wmatmul<DT>(d, lane, X4, E, Uacc);  // X4 is free once the odd polynomial has been consumed
      double2* t = Uacc;
      Uacc = X4;
      X4 = t;
      pos += 2;
    }
    pos += 1;
  }
  // out = V U V^+ (frame basis -> lab basis)
  double2* o = reinterpret_cast<double2*>(a.out) + b * (long long)n2;
  if (a.identityU) {
    for (int e = lane; e < n2; e += 32) o[e] = Uacc[e];
  } else {
    for (int e = lane; e < n2; e += 32) {  // X1 = V U
      const int i = e / d, j = e - i * d;
      double2 acc = make_double2(0.0, 0.0);
      for (int k = 0; k < d; k++) acc = cfma(__ldg(a.U + i * d + k), Uacc[k * d + j], acc);
      X1[e] = acc;
    }
    __syncwarp();
    for (int e = lane; e < n2; e += 32) {
      const int i = e / d, j = e - i * d;
      double2 acc = make_double2(0.0, 0.0);
      for (int k = 0; k < d; k++) {
        const double2 v = __ldg(a.U + j * d + k);
        acc = cfma(X1[i * d + k], make_double2(v.x, -v.y), acc);
      }
      o[e] = acc;
    }
  }
}

#include "propagator_block.cuh"

__global__ void sv_general_table_kernel(const double* __restrict__ times, long long NT, int KA, int d, int W,
                                        const double* __restrict__ two_pi_nu, const double* __restrict__ fe,
                                        double* __restrict__ out);
int casq_prepare_slots(const casq_model* m, int n_slots, const casq_pulse_slot* slots, DevSlots* ds, int* active,
                       int* n_active, int* n_total);

extern "C" int casq_propagators_expm(casq_model* m, int n_slots, const casq_pulse_slot* slots, int64_t B,
                                     const double* params_dev, double t0, double t1, double max_dt, double* out_dev,
                                     void* stream) {
  if (!m) CASQ_FAIL(CASQ_ERR_INVALID, "casq_propagators_expm: NULL model");
  if (B < 0) CASQ_FAIL(CASQ_ERR_INVALID, "casq_propagators_expm: negative batch size");
  if (B > 0 && (!out_dev || (n_slots > 0 && !params_dev))) CASQ_FAIL(CASQ_ERR_INVALID, "casq_propagators_expm: NULL buffer");
  if (!(t1 > t0)) CASQ_FAIL(CASQ_ERR_INVALID, "casq_propagators_expm: need t1 > t0");
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
  const int d = m->d;
  if (d < 2 || d > PROP_MAXD || KA > PROP_MAXK)
    CASQ_FAIL(CASQ_ERR_UNSUPPORTED, "casq_propagators_expm: dim %d with %d active channels is outside the built kernels (dim <= 64)", d, KA);
  cudaStream_t st = (cudaStream_t)stream;
  CASQ_CUDA(cudaSetDevice(m->device));
  TimeGrid g;
  rc = casq_build_time_grid(t0, t1, max_dt, 2, nullptr, m->dt, n_total, &g);
  if (rc != CASQ_OK) return rc;
  const long long NT = (long long)g.times.size();
  const int n_pieces = (int)g.piece_nsteps.size();
  const int W = 2 * KA + 2 * d, n2 = d * d;
  m->tab_key.clear();
  std::vector<double> consts(KA + d);
  for (int k = 0; k < KA; k++) consts[k] = 2.0 * M_PI * m->nu[active[k]];
  for (int i = 0; i < d; i++) consts[KA + i] = m->fe[i];
  std::vector<cplx> ops((size_t)(1 + 2 * KA) * n2 + n2);
  for (int e = 0; e < n2; e++) ops[e] = m->S[e];
  for (int k = 0; k < KA; k++)
    for (int e = 0; e < n2; e++) {
      cplx p = m->P[(size_t)active[k] * n2 + e];
      ops[(size_t)(1 + k) * n2 + e] = m->rwa ? p : 2.0 * p;
      ops[(size_t)(1 + KA + k) * n2 + e] = m->N[(size_t)active[k] * n2 + e];
    }
  for (int e = 0; e < n2; e++) ops[(size_t)(1 + 2 * KA) * n2 + e] = m->U[e];
  if ((rc = m->tab_times.ensure(NT * sizeof(double)))) return rc;
  if ((rc = m->tab_idx.ensure(NT * sizeof(int)))) return rc;
  if ((rc = m->tab.ensure((size_t)NT * W * sizeof(double)))) return rc;
  if ((rc = m->pieces_n.ensure(n_pieces * sizeof(int)))) return rc;
  if ((rc = m->pieces_h.ensure(n_pieces * sizeof(double)))) return rc;
  if ((rc = m->work0.ensure(consts.size() * sizeof(double)))) return rc;
  if ((rc = m->work1.ensure(ops.size() * sizeof(cplx)))) return rc;
  CASQ_CUDA(cudaMemcpyAsync(m->tab_times.p, g.times.data(), NT * sizeof(double), cudaMemcpyHostToDevice, st));
  CASQ_CUDA(cudaMemcpyAsync(m->tab_idx.p, g.sample_idx.data(), NT * sizeof(int), cudaMemcpyHostToDevice, st));
  CASQ_CUDA(cudaMemcpyAsync(m->pieces_n.p, g.piece_nsteps.data(), n_pieces * sizeof(int), cudaMemcpyHostToDevice, st));
  CASQ_CUDA(cudaMemcpyAsync(m->pieces_h.p, g.piece_h.data(), n_pieces * sizeof(double), cudaMemcpyHostToDevice, st));
  CASQ_CUDA(cudaMemcpyAsync(m->work0.p, consts.data(), consts.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  CASQ_CUDA(cudaMemcpyAsync(m->work1.p, ops.data(), ops.size() * sizeof(cplx), cudaMemcpyHostToDevice, st));
  {
    const long long nthreads = NT * (KA + d);
    sv_general_table_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>(
        (const double*)m->tab_times.p, NT, KA, d, W, (const double*)m->work0.p, (const double*)m->work0.p + KA,
        (double*)m->tab.p);
    CASQ_CUDA(cudaGetLastError());
  }
  PropArgs a;
  memset(&a, 0, sizeof(a));
  a.table = (const double*)m->tab.p;
  a.idx = (const int*)m->tab_idx.p;
  a.piece_n = (const int*)m->pieces_n.p;
  a.piece_h = (const double*)m->pieces_h.p;
  a.n_pieces = n_pieces;
  a.B = B;
  a.d = d;
  a.KA = KA;
  a.W = W;
  a.rwa = m->rwa ? 1 : 0;
  a.params = params_dev;
  a.out = out_dev;
  a.slots = ds;
  a.S = (const double2*)m->work1.p;
  a.P = a.S + n2;
  a.N = a.P + (size_t)KA * n2;
  a.U = a.N + (size_t)KA * n2;
  a.identityU = m->U_is_identity ? 1 : 0;
  const int wpb = d <= 16 ? PROP_WARPS : 2;
  const size_t smem = (size_t)wpb * PROP_NBUF * n2 * sizeof(double2);
  // the batch is walked in windows that bound the envelope table
  const long long Bw = casq_env_table_window(KA, n_total, B);
  if ((rc = m->work2.ensure((size_t)KA * std::max(n_total, 1) * (size_t)Bw * sizeof(double2)))) return rc;
  a.env_tab = (const double2*)m->work2.p;
  a.n_total = n_total;
  // the dimensions transmon models produce get compile-time loop bounds (index arithmetic was 80 % of the generic kernel)
#define PROP_LAUNCH(DT_)                                                                                              \
  {                                                                                                                   \
    CASQ_CUDA(cudaFuncSetAttribute(propagator_expm_kernel<DT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    propagator_expm_kernel<DT_><<<grid, 32 * wpb, smem, st>>>(a);                                                     \
  }
  for (long long b0 = 0; b0 < B; b0 += Bw) {
    const long long Bc = std::min(Bw, (long long)B - b0);
    CASQ_CUDA(casq_launch_env_table<PROP_MAXK>(KA, ds, params_dev, B, b0, Bc, n_total, (double2*)m->work2.p, st));
    a.B = Bc;
    a.out = out_dev + b0 * 2LL * n2;
    if (d > 16) {  // a block per trajectory (propagator_block.cuh); work matrices in shared memory up to dim 32
      size_t smem_b = (size_t)PROP_NBUF * n2 * sizeof(double2);
      a.scratch = nullptr;
      if (d > 32) {
        if ((rc = m->work3.ensure((size_t)Bc * smem_b))) return rc;
        a.scratch = (double2*)m->work3.p;
        smem_b = 0;
      }
#define PROP_BLOCK_LAUNCH(TS_)                                                                                                             \
  {                                                                                                                                        \
    if (a.scratch) {                                                                                                                       \
      propagator_expm_block_kernel<TS_, true><<<(unsigned)Bc, PB_THREADS, 0, st>>>(a);                                                     \
    } else {                                                                                                                               \
      CASQ_CUDA(cudaFuncSetAttribute(propagator_expm_block_kernel<TS_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b)); \
      propagator_expm_block_kernel<TS_, false><<<(unsigned)Bc, PB_THREADS, smem_b, st>>>(a);                                               \
    }                                                                                                                                      \
  }
      // products on the FP64 tensor pipe (DMMA split-complex GEMM) unless CASQ_PROP_NO_MMA asks for the CUDA-core tiles
      if (!getenv("CASQ_PROP_NO_MMA")) PROP_BLOCK_LAUNCH(0)
      else if (d % 3 == 0) PROP_BLOCK_LAUNCH(3)
      else if (d % 2 == 0) PROP_BLOCK_LAUNCH(2)
      else PROP_BLOCK_LAUNCH(1)
#undef PROP_BLOCK_LAUNCH
      continue;
    }
    const unsigned grid = (unsigned)((Bc + wpb - 1) / wpb);
    switch (d) {
      case 2: PROP_LAUNCH(2) break;
      case 3: PROP_LAUNCH(3) break;
      case 4: PROP_LAUNCH(4) break;
      case 9: PROP_LAUNCH(9) break;
      default: PROP_LAUNCH(0) break;
    }
  }
#undef PROP_LAUNCH
  CASQ_CUDA(cudaGetLastError());
  return CASQ_OK;
}
