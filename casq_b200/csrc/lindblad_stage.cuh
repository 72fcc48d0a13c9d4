// casq_b200: fused Lindblad RK4 stage kernel ("tile pairs"), the hot kernel of BASELINE cfg 4.
//
// One block owns the 64x64 tile pair (I,J), (J,I), I <= J, of one trajectory's density matrix:
//   1. M_JI = (A rho)[J,I]  (left gathers only)                                   -> shared tile s_T
//   2. M_IJ = (A rho)[I,J] + (two-sided + dense-diagonal terms)[I,J], row by row; k = M_IJ + M_JI^+ (read transposed
//      from s_T; rho Hermitian => rho A^+ = (A rho)^+); RK4 combination with coalesced y / out accesses; the result
//      overwrites the s_T slot it consumed
//   3. the mirror tile out[J,I] = out[I,J]^+ is written from s_T: y[J,I], y1[J,I], ... are never read and every stage
//      output is exactly Hermitian.
//   Diagonal tiles (I == J) compute M_II with HALF the two-sided terms once and form k = M_II + M_II^+ from s_T.
// Against the SpMM + combine pair it replaces (profiles/r01_ncu_lindblad_v6.txt): no M round trip through HBM
// (~11.4 instead of 27 matrix passes per step), two-sided terms on the upper triangle only, one launch per stage.
// Profiles of the intermediate versions showed the path is bound by ISSUED INSTRUCTIONS at IPC ~0.45 (in-order
// chains of address arithmetic around each 512 B gather), so each lane owns TWO columns (c, c + 32): one list
// look-up and one address computation feed two LDG.128 and eight DFMA.
// grid (NT(NT+1)/2, B), 16 warps: warp w owns rows w, w+16, w+32, w+48 of each tile.
#pragma once
#include "lindblad_kernels.cuh"

#define LB_TILE 64
#define LB_TPAD 65

struct LbTileShared {  // block-wide group tables
  int gs[LB_MAXQ + 1];
  int gd1[LB_MAXQ], gd2[LB_MAXQ];
  int gp[4 * LB_MAXQ];
};

// One output row segment: columns c0 and c0 + 32 of row r.  Lists in shared memory (warp-private).
template <bool TWO>
__device__ __forceinline__ void lb_tile_unit(int d, int ld, const double2* lc, const int* lo, int cnt, const double2* ul, const int* sg, int gc,
                                             const LbTileShared& T, const double2* __restrict__ cdense,
                                             const double2* __restrict__ gu, const double2* __restrict__ in, int r, int c0, bool ok0,
                                             bool ok1, double w, double2& o0, double2& o1) {
  const int e0 = r * ld + c0;  // d <= 1024: element indices of one matrix fit 32 bits
  const double2* pc = in + e0;
  asm volatile("" : "+l"(pc));  // keep the lane's base pointer in registers: one IMAD.WIDE per gather
  const double2 z = make_double2(0.0, 0.0);
  double ax = 0.0, ay = 0.0, bx = 0.0, by = 0.0;
#pragma unroll 2
  for (int k = 0; k < cnt; k++) {
    const double2 cf = lc[k];
    const double2* p = pc + lo[k];
    const double2 v0 = ok0 ? __ldg(p) : z;
    const double2 v1 = ok1 ? __ldg(p + 32) : z;
    ax = fma(cf.x, v0.x, fma(-cf.y, v0.y, ax));
    ay = fma(cf.x, v0.y, fma(cf.y, v0.x, ay));
    bx = fma(cf.x, v1.x, fma(-cf.y, v1.y, bx));
    by = fma(cf.x, v1.y, fma(cf.y, v1.x, by));
  }
  if (TWO) {
    if (cdense) {
      const double w2 = 2.0 * w;  // stored x 1/2
      if (ok0) {
        const double2 cf = __ldg(cdense + e0), v = __ldg(pc);
        ax = fma(w2 * cf.x, v.x, fma(-w2 * cf.y, v.y, ax));
        ay = fma(w2 * cf.x, v.y, fma(w2 * cf.y, v.x, ay));
      }
      if (ok1) {
        const double2 cf = __ldg(cdense + e0 + 32), v = __ldg(pc + 32);
        bx = fma(w2 * cf.x, v.x, fma(-w2 * cf.y, v.y, bx));
        by = fma(w2 * cf.x, v.y, fma(w2 * cf.y, v.x, by));
      }
    }
    for (int kk = 0; kk < gc; kk++) {
      // A group acts on this row only if r + delta1 is a valid row; shifted columns outside [0,d) carry a zero
      // coefficient (their right factor is zero) and are not loaded.
      const int g = sg[kk];
      const int cc0 = c0 + T.gd2[g], cc1 = cc0 + 32;
      const double2* prow = in + (r + T.gd1[g]) * ld;
      const double2 v0 = (ok0 && (unsigned)cc0 < (unsigned)d) ? __ldg(prow + cc0) : z;
      const double2 v1 = (ok1 && (unsigned)cc1 < (unsigned)d) ? __ldg(prow + cc1) : z;
      double g0x = 0.0, g0y = 0.0, g1x = 0.0, g1y = 0.0;
      for (int p = T.gs[g]; p < T.gs[g + 1]; p++) {
        const double2 l = ul[T.gp[2 * p]];
        const double2* ur = gu + T.gp[2 * p + 1] * d + c0;
        const double2 r0 = ok0 ? __ldg(ur) : z, r1 = ok1 ? __ldg(ur + 32) : z;  // conjugated below
        g0x = fma(l.x, r0.x, fma(l.y, r0.y, g0x));
        g0y = fma(l.y, r0.x, fma(-l.x, r0.y, g0y));
        g1x = fma(l.x, r1.x, fma(l.y, r1.y, g1x));
        g1y = fma(l.y, r1.x, fma(-l.x, r1.y, g1y));
      }
      ax = fma(g0x, v0.x, fma(-g0y, v0.y, ax));
      ay = fma(g0x, v0.y, fma(g0y, v0.x, ay));
      bx = fma(g1x, v1.x, fma(-g1y, v1.y, bx));
      by = fma(g1x, v1.y, fma(g1y, v1.x, by));
    }
  }
  o0 = make_double2(ax, ay);
  o1 = make_double2(bx, by);
}

__device__ __forceinline__ double2 lb_rk4_combine(const LbStage& s, long long o, double2 y0, double kx, double ky) {
  if (s.stage == 1 || s.stage == 2) return make_double2(fma(0.5 * s.h, kx, y0.x), fma(0.5 * s.h, ky, y0.y));
  if (s.stage == 3) return make_double2(fma(s.h, kx, y0.x), fma(s.h, ky, y0.y));
  const double2 v1 = s.y1[o], v2 = s.y2[o], v3 = s.in[o];
  const double t3 = 1.0 / 3.0, h6 = s.h * (1.0 / 6.0);
  return make_double2(fma(h6, kx, t3 * (v1.x + 2.0 * v2.x + v3.x - y0.x)), fma(h6, ky, t3 * (v1.y + 2.0 * v2.y + v3.y - y0.y)));
}

__global__ void __launch_bounds__(512, 2) lindblad_stage_kernel(const LbArgs a, const LbLists L, const double2* __restrict__ cu, int set,
                                                                const LbStage s) {
  extern __shared__ __align__(16) unsigned char lb_smem[];
  __shared__ LbTileShared T;
  const int d = a.d, ld = a.ld, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int NQs = (a.NQ + 3) & ~3, NUs = a.NU > 0 ? a.NU : 1, NGs = a.NG > 0 ? a.NG : 1;
  double2(*s_T)[LB_TPAD] = reinterpret_cast<double2(*)[LB_TPAD]>(lb_smem);          // [64][65]
  double2* dbase = reinterpret_cast<double2*>(lb_smem) + LB_TILE * LB_TPAD;
  double2* lc = dbase + (size_t)warp * 8 * NQs;                                     // [16 warps][8 rows][NQs]
  double2* ul = dbase + (size_t)16 * 8 * NQs + (size_t)warp * 4 * NUs;              // [16][4][NUs]
  int* ibase = reinterpret_cast<int*>(dbase + (size_t)16 * 8 * NQs + (size_t)16 * 4 * NUs);
  int* lo = ibase + (size_t)warp * 8 * NQs;                                         // [16][8][NQs]
  int* sg = ibase + (size_t)16 * 8 * NQs + (size_t)warp * 4 * NGs;                  // [16][4][NGs]
  const long long b = blockIdx.y;
  const int I = a.pair_ij[2 * blockIdx.x], J = a.pair_ij[2 * blockIdx.x + 1];
  const int rI = LB_TILE * I, rJ = LB_TILE * J;
  const bool diag = I == J;
  if (threadIdx.x <= a.NG) T.gs[threadIdx.x] = a.grp_start[threadIdx.x];
  if (threadIdx.x < 2 * a.NP) T.gp[threadIdx.x] = a.grp_pairs[threadIdx.x];
  if (threadIdx.x < a.NG) {
    const int p0 = a.grp_start[threadIdx.x];
    T.gd1[threadIdx.x] = a.u_delta[a.grp_pairs[2 * p0]];
    T.gd2[threadIdx.x] = a.u_delta[a.grp_pairs[2 * p0 + 1]];
  }
  const long long base = b * (long long)d * ld;
  const double2* in = s.in + base;
  const double2* gu = cu + ((long long)set * a.B + b) * a.NU * d;
  const long long rowbase = ((long long)set * a.B + b) * d;
  const double w = diag ? 0.5 : 1.0;
  // list lengths, one per lane: lanes 0-3 rows of the I band, 4-7 rows of the J band, 8-11 group counts of the I band
  int mycnt = 0;
  if (lane < 12) {
    const int r = ((lane & 12) == 4 ? rJ : rI) + warp + 16 * (lane & 3);
    if (r < d && ((lane & 12) != 4 || !diag)) mycnt = lane < 8 ? L.lcnt[rowbase + r] : L.gcnt[rowbase + r];
  }
  // per-row lists of the 8 rows this warp touches: loads first (rows clamped, so they are unconditional), then stores
  const int q = lane < a.NQ ? lane : 0;
#pragma unroll
  for (int half = 0; half < 2; half++) {
    double2 tc[4];
    int to[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int r = (half ? rJ : rI) + warp + 16 * u;
      const long long rowid = rowbase + (r < d ? r : d - 1);
      tc[u] = L.lcoef[rowid * a.NQ + q];
      to[u] = L.loff[rowid * a.NQ + q];
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      lc[(4 * half + u) * NQs + q] = tc[u];  // entries >= the row's count are stale and never used
      lo[(4 * half + u) * NQs + q] = to[u];
    }
  }
#pragma unroll
  for (int u = 0; u < 4; u++) {
    const int r = rI + warp + 16 * u, rc = r < d ? r : d - 1;
    const long long rowid = rowbase + rc;
    if (lane < a.NG) sg[u * NGs + lane] = L.glist[rowid * a.NG + lane];
    if (lane < a.NU) {
      const double2 v = gu[lane * d + rc];
      ul[u * NUs + lane] = make_double2(w * v.x, w * v.y);
    }
  }
  __syncthreads();
  const int cI = rI + lane, cJ = rJ + lane;
  const bool okI0 = cI < d, okI1 = cI + 32 < d, okJ0 = cJ < d, okJ1 = cJ + 32 < d;
  // ---- 1. M_JI (or, on the diagonal, M_II with half the two-sided terms) into s_T
#pragma unroll 1
  for (int u = 0; u < 4; u++) {
    const int x = warp + 16 * u, r = rJ + x;
    const int cnt = __shfl_sync(0xffffffffu, mycnt, diag ? u : 4 + u), gc = __shfl_sync(0xffffffffu, mycnt, 8 + u);
    double2 o0 = make_double2(0.0, 0.0), o1 = o0;
    if (r < d) {
      if (diag)
        lb_tile_unit<true>(d, ld, lc + u * NQs, lo + u * NQs, cnt, ul + u * NUs, sg + u * NGs, gc, T, a.cdense, gu, in, r, cI, okI0, okI1, w, o0, o1);
      else
        lb_tile_unit<false>(d, ld, lc + (4 + u) * NQs, lo + (4 + u) * NQs, cnt, nullptr, nullptr, 0, T, nullptr, gu, in, r, cI, okI0, okI1, w, o0, o1);
    }
    s_T[x][lane] = o0;
    s_T[x][lane + 32] = o1;
  }
  __syncthreads();
  // ---- 2. rows of tile (I,J): k, RK4 combination, store; the result replaces the consumed s_T slot
#pragma unroll 1
  for (int u = 0; u < 4; u++) {
    const int x = warp + 16 * u, r = rI + x;
    const int cnt = __shfl_sync(0xffffffffu, mycnt, u), gc = __shfl_sync(0xffffffffu, mycnt, 8 + u);
    if (r >= d) continue;  // warp-uniform
    // step-start values first: their HBM latency overlaps the gathers below
    const long long o = base + (long long)r * ld + cJ;
    const double2 zz = make_double2(0.0, 0.0);
    const double2 y00 = okJ0 ? s.y[o] : zz, y01 = okJ1 ? s.y[o + 32] : zz;
    double2 o0, o1;
    if (diag) {
      o0 = s_T[x][lane];
      o1 = s_T[x][lane + 32];
    } else {
      lb_tile_unit<true>(d, ld, lc + u * NQs, lo + u * NQs, cnt, ul + u * NUs, sg + u * NGs, gc, T, a.cdense, gu, in, r, cJ, okJ0, okJ1, w, o0, o1);
    }
    if (okJ0) {
      const double2 t = s_T[lane][x];  // M_JI(rJ + lane, rI + x)
      const double2 res = lb_rk4_combine(s, o, y00, o0.x + t.x, o0.y - t.y);
      s.out[o] = res;
      if (!diag) s_T[lane][x] = res;
    }
    if (okJ1) {
      const double2 t = s_T[lane + 32][x];
      const double2 res = lb_rk4_combine(s, o + 32, y01, o1.x + t.x, o1.y - t.y);
      s.out[o + 32] = res;
      if (!diag) s_T[lane + 32][x] = res;
    }
  }
  if (diag) return;
  __syncthreads();
  // ---- 3. mirror tile: out(rJ + x, rI + y) = conj(out(rI + y, rJ + x)) = conj(s_T[x][y])
#pragma unroll
  for (int u = 0; u < 4; u++) {
    const int x = warp + 16 * u, r = rJ + x;
    if (r < d) {
      const long long o = base + (long long)r * ld + cI;
      if (okI0) {
        const double2 t = s_T[x][lane];
        s.out[o] = make_double2(t.x, -t.y);
      }
      if (okI1) {
        const double2 t = s_T[x][lane + 32];
        s.out[o + 32] = make_double2(t.x, -t.y);
      }
    }
  }
}

static inline size_t lindblad_stage_smem(int NQ, int NU, int NG) {
  const size_t NQs = (size_t)((NQ + 3) & ~3), NUs = (size_t)(NU > 0 ? NU : 1), NGs = (size_t)(NG > 0 ? NG : 1);
  return (size_t)LB_TILE * LB_TPAD * 16 + 16 * 8 * NQs * 16 + 16 * 4 * NUs * 16 + 16 * 8 * NQs * 4 + 16 * 4 * NGs * 4;
}
