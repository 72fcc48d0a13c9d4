// casq_b200: propagators for 16 < dim <= 64 (three coupled transmons: dim 27), ONE BLOCK of 256 threads per trajectory.
//
// Six dim x dim work matrices are 70 kB at dim 27: with a warp per trajectory only two warps fit an SM and every
// lane walks 23 matrix elements per product (the generic kernel: 1.0e3 unitaries/s at dim 27).  Here the 256 threads
// of a block share one trajectory: products are register tiled (a thread owns a TS x TS output tile, TS = 3 when
// 3 | dim, else 2 or 1: per k it loads 2 TS operands for TS^2 complex FMAs), the Gauss-Jordan elimination and all
// element-wise passes are strided over the block, and three blocks (24 warps) are resident per SM.
// Same algorithm as propagator.cu (orders 3/5/7/9 by Higham's theta_m, scaling above theta_9, partial pivoting).
// Up to dim 32 the six work matrices live in shared memory; for 32 < dim <= 64 (393 kB at dim 64) they live in a
// per-trajectory global scratch that stays L2 resident, reached through the same pointers (__syncthreads orders the
// block's global writes and reads).
#pragma once

#define PB_THREADS 256

// C = X * Y (d x d), block-cooperative; C must not alias X or Y.  A thread owns a TR x TS output tile.
#ifndef PB_TILE_ROWS
#define PB_TILE_ROWS(TS) (TS)  // square tiles: 1 x TS tiles put 3x the threads to work at dim 27 but ran 11 % slower (operand traffic)
#endif
// ---- tensor-core product: split-complex GEMM on the FP64 tensor pipe (DMMA.8x8x4 = mma.sync m8n8k4 f64) -------------------
// north_star (3): "run on tensor cores via split-complex GEMM only where d >= 16 makes it a real dense contraction".
// C = X Y with X = Xr + i Xi, Y = Yr + i Yi as four real 8x8x4 MMAs per k-step: Cr += Xr Yr, Cr += (-Xi) Yi, Ci += Xr Yi,
// Ci += Xi Yr, accumulators in registers (fp64: same precision as the FMA path, results differ by summation order only).
// A warp owns an 8 x 16 strip of C (two 8 x 8 tiles sharing the X fragment); fragments are read straight from the
// interleaved complex matrices: one 16 B load gives the (re, im) pair of an operand element, so a k-step of 4 costs 3
// loads per lane for 8 MMAs = 512 complex FMAs per warp (the register-tiled CUDA-core product loads 6 operands per 9).
// m8n8k4 fragment layout: a = A[lane / 4][lane % 4], b = B[lane % 4][lane / 4], c0/c1 = C[lane / 4][2 (lane % 4) + 0/1].
__device__ __forceinline__ void pb_dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

__device__ __forceinline__ void bmatmul_mma(int d, int tid, double2* C, const double2* X, const double2* Y) {
  const int lane = tid & 31, warp = tid >> 5, gr = lane >> 2, gk = lane & 3;
  const int nt = (d + 7) >> 3, npair = (nt + 1) >> 1;
  const double2 z = make_double2(0.0, 0.0);
  for (int job = warp; job < nt * npair; job += PB_THREADS / 32) {
    const int ti = job / npair, tp = job - ti * npair;
    const int r = ti * 8 + gr, ca = tp * 16 + gr, cb = ca + 8;
    const bool rok = r < d, caok = ca < d, cbok = cb < d;
    double cr0[2] = {0.0, 0.0}, ci0[2] = {0.0, 0.0}, cr1[2] = {0.0, 0.0}, ci1[2] = {0.0, 0.0};
    const double2* xp = X + r * d + gk;
    const double2* yp = Y + gk * d + ca;
#pragma unroll 2
    for (int k0 = 0; k0 < d; k0 += 4) {
      const bool kok = k0 + gk < d;
      const double2 a = (rok && kok) ? xp[k0] : z;
      const double2 b0 = (kok && caok) ? yp[k0 * d] : z;
      const double2 b1 = (kok && cbok) ? yp[k0 * d + 8] : z;
      const double nai = -a.y;
      pb_dmma(cr0, a.x, b0.x);
      pb_dmma(ci0, a.x, b0.y);
      pb_dmma(cr1, a.x, b1.x);
      pb_dmma(ci1, a.x, b1.y);
      pb_dmma(cr0, nai, b0.y);
      pb_dmma(ci0, a.y, b0.x);
      pb_dmma(cr1, nai, b1.y);
      pb_dmma(ci1, a.y, b1.x);
    }
    if (rok) {
      const int c0 = tp * 16 + 2 * gk;
      double2* o = C + r * d + c0;
      if (c0 < d) o[0] = make_double2(cr0[0], ci0[0]);
      if (c0 + 1 < d) o[1] = make_double2(cr0[1], ci0[1]);
      if (c0 + 8 < d) o[8] = make_double2(cr1[0], ci1[0]);
      if (c0 + 9 < d) o[9] = make_double2(cr1[1], ci1[1]);
    }
  }
  __syncthreads();
}

// TS = 0 selects the tensor-core product above; TS >= 1 the register-tiled CUDA-core product (kept for A/B runs:
// CASQ_PROP_NO_MMA=1).
template <int TS>
__device__ __forceinline__ void bmatmul(int d, int tid, double2* C, const double2* X, const double2* Y) {
  if constexpr (TS == 0) {
    bmatmul_mma(d, tid, C, X, Y);
    return;
  }
  constexpr int TR = PB_TILE_ROWS(TS == 0 ? 1 : TS);
  const int ntr = d / TR, ntc = d / (TS == 0 ? 1 : TS), ntiles = ntr * ntc;
  for (int tile = tid; tile < ntiles; tile += PB_THREADS) {
    const int ti = tile / ntc, r0 = ti * TR, c0 = (tile - ti * ntc) * TS;
    constexpr int TC = TS == 0 ? 1 : TS;
    double2 acc[TR][TC];
#pragma unroll
    for (int i = 0; i < TR; i++)
#pragma unroll
      for (int j = 0; j < TS; j++) acc[i][j] = make_double2(0.0, 0.0);
    for (int k = 0; k < d; k++) {
      double2 x[TR], y[TC];
#pragma unroll
      for (int i = 0; i < TR; i++) x[i] = X[(r0 + i) * d + k];
#pragma unroll
      for (int j = 0; j < TS; j++) y[j] = Y[k * d + c0 + j];
#pragma unroll
      for (int i = 0; i < TR; i++)
#pragma unroll
        for (int j = 0; j < TS; j++) acc[i][j] = cfma(x[i], y[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < TR; i++)
#pragma unroll
      for (int j = 0; j < TS; j++) C[(r0 + i) * d + c0 + j] = acc[i][j];
  }
  __syncthreads();
}

// Solve Q X = Pm in place (X overwrites Pm); Q is destroyed.  Gauss-Jordan with partial pivoting; the pivot search is
// done by warp 0 (d <= 64: at most two rows per lane) and the chosen row travels through s_piv.  Pivot rows are NOT normalised inside
// the loop (the factor carries 1/pivot; every row is divided by its own pivot once at the end): two block barriers
// per pivot instead of four.  thread = (row, column group) in the elimination: the factor is computed once per thread.
template <int ROWS>  // 32 or 64: padded row count (d <= ROWS)
__device__ __forceinline__ void bsolve(int d, int tid, double2* Q, double2* Pm, int* s_piv) {
  constexpr int CG = PB_THREADS / ROWS;
  const int r = tid / CG, jg = tid - r * CG;
  for (int k = 0; k < d; k++) {
    if (tid < 32) {
      double best = -1.0;
      int brow = k;
      for (int rr = k + tid; rr < d; rr += 32) {
        const double2 v = Q[rr * d + k];
        const double mag = v.x * v.x + v.y * v.y;
        if (mag > best) {
          best = mag;
          brow = rr;
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
      if (tid == 0) *s_piv = brow;
    }
    __syncthreads();
    const int brow = *s_piv;
    if (brow != k) {  // block-uniform
      for (int j = tid; j < 2 * d; j += PB_THREADS) {
        double2* M = (j < d) ? Q : Pm;
        const int c = (j < d) ? j : j - d;
        const double2 t = M[k * d + c];
        M[k * d + c] = M[brow * d + c];
        M[brow * d + c] = t;
      }
      __syncthreads();
    }
    // eliminate column k from every other row.  Columns <= k of Q are dead (never written), so the pivot and the
    // factors stay readable while the live columns are updated; row k itself is not touched.
    if (r < d && r != k) {
      const double2 piv = Q[k * d + k];
      const double inv_n = 1.0 / (piv.x * piv.x + piv.y * piv.y);
      const double2 f = cmul(Q[r * d + k], make_double2(piv.x * inv_n, -piv.y * inv_n));
      for (int j = k + 1 + jg; j < 2 * d; j += CG) {
        double2* M = (j < d) ? Q : Pm;
        const int c = (j < d) ? j : j - d;
        const double2 p = M[k * d + c];
        const double2 v = M[r * d + c];
        M[r * d + c] = make_double2(fma(-f.x, p.x, fma(f.y, p.y, v.x)), fma(-f.x, p.y, fma(-f.y, p.x, v.y)));
      }
    }
    __syncthreads();
  }
  // X[r][:] = P[r][:] / Q[r][r]
  if (r < d) {
    const double2 piv = Q[r * d + r];
    const double inv_n = 1.0 / (piv.x * piv.x + piv.y * piv.y);
    const double2 pinv = make_double2(piv.x * inv_n, -piv.y * inv_n);
    for (int c = jg; c < d; c += CG) Pm[r * d + c] = cmul(Pm[r * d + c], pinv);
  }
  __syncthreads();
}

// expm(A) with four work matrices; returns the buffer holding the result (X1 or X2).  See wexpm (propagator.cu).
template <int TS, int ROWS>
__device__ double2* bexpm(int d, int tid, double2* A, double2* X1, double2* X2, double2* X3, double2* X4, double* s_red, int* s_piv) {
  const int n2 = d * d;
  for (int e = tid; e < n2; e += PB_THREADS) X4[e].x = sqrt(A[e].x * A[e].x + A[e].y * A[e].y);
  __syncthreads();
  if (tid < 32) {  // ||A||_1: lane j sums columns j, j + 32 (d <= 64), warp max
    double s = 0.0;
    for (int j = tid; j < d; j += 32) {
      double cs = 0.0;
      for (int i = 0; i < d; i++) cs += X4[i * d + j].x;
      s = fmax(s, cs);
    }
    for (int o = 16; o > 0; o >>= 1) s = fmax(s, __shfl_xor_sync(0xffffffffu, s, o));
    if (tid == 0) *s_red = s;
  }
  __syncthreads();
  const double nrm = *s_red;
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
      for (int e = tid; e < n2; e += PB_THREADS) A[e] = make_double2(A[e].x * sc, A[e].y * sc);
    }
  }
  __syncthreads();  // also: every thread has read s_red / finished with the |A| scratch
  double2 *A2 = X1, *A4 = X2, *A6 = X3, *A8 = X4, *Od = X4, *Vm = X3;
  bmatmul<TS>(d, tid, A2, A, A);
  auto diag = [&](int e) { return (e / d) == (e % d) ? 1.0 : 0.0; };
  if (m == 3) {
    const double b[4] = {120., 60., 12., 1.};
    for (int e = tid; e < n2; e += PB_THREADS) {
      Od[e] = make_double2(b[3] * A2[e].x + b[1] * diag(e), b[3] * A2[e].y);
      Vm[e] = make_double2(b[2] * A2[e].x + b[0] * diag(e), b[2] * A2[e].y);
    }
  } else if (m == 5) {
    const double b[6] = {30240., 15120., 3360., 420., 30., 1.};
    bmatmul<TS>(d, tid, A4, A2, A2);
    for (int e = tid; e < n2; e += PB_THREADS) {
      Od[e] = make_double2(b[5] * A4[e].x + b[3] * A2[e].x + b[1] * diag(e), b[5] * A4[e].y + b[3] * A2[e].y);
      Vm[e] = make_double2(b[4] * A4[e].x + b[2] * A2[e].x + b[0] * diag(e), b[4] * A4[e].y + b[2] * A2[e].y);
    }
  } else if (m == 7) {
    const double b[8] = {17297280., 8648640., 1995840., 277200., 25200., 1512., 56., 1.};
    bmatmul<TS>(d, tid, A4, A2, A2);
    bmatmul<TS>(d, tid, A6, A4, A2);
    for (int e = tid; e < n2; e += PB_THREADS) {  // V overwrites A^6 element by element
      const double2 a6 = A6[e];
      Od[e] = make_double2(b[7] * a6.x + b[5] * A4[e].x + b[3] * A2[e].x + b[1] * diag(e), b[7] * a6.y + b[5] * A4[e].y + b[3] * A2[e].y);
      Vm[e] = make_double2(b[6] * a6.x + b[4] * A4[e].x + b[2] * A2[e].x + b[0] * diag(e), b[6] * a6.y + b[4] * A4[e].y + b[2] * A2[e].y);
    }
  } else {
    const double b[10] = {17643225600., 8821612800., 2075673600., 302702400., 30270240., 2162160., 110880., 3960., 90., 1.};
    bmatmul<TS>(d, tid, A4, A2, A2);
    bmatmul<TS>(d, tid, A6, A4, A2);
    bmatmul<TS>(d, tid, A8, A6, A2);
    for (int e = tid; e < n2; e += PB_THREADS) {  // the polynomials overwrite A^8 and A^6 element by element
      const double2 a8 = A8[e], a6 = A6[e];
      Od[e] = make_double2(b[9] * a8.x + b[7] * a6.x + b[5] * A4[e].x + b[3] * A2[e].x + b[1] * diag(e),
                           b[9] * a8.y + b[7] * a6.y + b[5] * A4[e].y + b[3] * A2[e].y);
      Vm[e] = make_double2(b[8] * a8.x + b[6] * a6.x + b[4] * A4[e].x + b[2] * A2[e].x + b[0] * diag(e),
                           b[8] * a8.y + b[6] * a6.y + b[4] * A4[e].y + b[2] * A2[e].y);
    }
  }
  __syncthreads();
  double2* Um = X1;  // A^2 is dead
  bmatmul<TS>(d, tid, Um, A, Od);
  for (int e = tid; e < n2; e += PB_THREADS) {  // Q = V - U (over V), P = V + U (over U)
    const double2 v = Vm[e], u = Um[e];
    X3[e] = make_double2(v.x - u.x, v.y - u.y);
    X1[e] = make_double2(v.x + u.x, v.y + u.y);
  }
  __syncthreads();
  bsolve<ROWS>(d, tid, X3, X1, s_piv);
  double2 *E = X1, *Sq = X2;
  for (int q = 0; q < s; q++) {
    bmatmul<TS>(d, tid, Sq, E, E);
    double2* t = E;
    E = Sq;
    Sq = t;
  }
  return E;
}

// GWS: work matrices in the global scratch (dim > 32); a compile-time switch, so that in the shared-memory case the
// compiler still knows every matrix pointer is shared (generic loads cost 15 % at dim 27).
template <int TS, bool GWS>
__global__ void __launch_bounds__(PB_THREADS) propagator_expm_block_kernel(const PropArgs a) {
  const int d = a.d, n2 = d * d, tid = threadIdx.x;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double s_red;
  __shared__ int s_piv;
  const long long b = blockIdx.x;
  double2* ws = GWS ? a.scratch + (size_t)b * PROP_NBUF * n2 : reinterpret_cast<double2*>(smem_raw);
  double2 *A = ws, *Uacc = ws + n2, *X1 = ws + 2 * n2, *X2 = ws + 3 * n2, *X3 = ws + 4 * n2, *X4 = ws + 5 * n2;
  for (int e = tid; e < n2; e += PB_THREADS) Uacc[e] = make_double2((e / d) == (e % d) ? 1.0 : 0.0, 0.0);
  __syncthreads();

  long long pos = 0;
  for (int p = 0; p < a.n_pieces; p++) {
    const int n = a.piece_n[p];
    const double h = a.piece_h[p];
    for (int s = 0; s < n; s++) {
      const long long pm = pos + 1;  // midpoint entry
      double zr[PROP_MAXK], zi[PROP_MAXK];
      prop_signals(a, b, pm, zr, zi);
      prop_assemble<0>(a, d, pm, zr, zi, h, A, tid, PB_THREADS);
      __syncthreads();
      double2* E = bexpm<TS, (GWS ? 64 : 32)>(d, tid, A, X1, X2, X3, X4, &s_red, &s_piv);
      bmatmul<TS>(d, tid, X4, E, Uacc);  // X4 is free once the odd polynomial has been consumed
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
    for (int e = tid; e < n2; e += PB_THREADS) o[e] = Uacc[e];
  } else {
    for (int e = tid; e < n2; e += PB_THREADS) {  // X1 = V U
      const int i = e / d, j = e - i * d;
      double2 acc = make_double2(0.0, 0.0);
      for (int k = 0; k < d; k++) acc = cfma(__ldg(a.U + i * d + k), Uacc[k * d + j], acc);
      X1[e] = acc;
    }
    __syncthreads();
    for (int e = tid; e < n2; e += PB_THREADS) {
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
