// does a DFMA warp-instruction with only the low 16 (or 8) lanes active take half the fp64-pipe time?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void chain(int iters, int active_lanes, double* out, long long* cyc) {
  double a[4];
  for (int i = 0; i < 4; i++) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
  const double m = 1.0 - 1e-12, c = 1e-13;
  long long t0 = clock64();
  if ((threadIdx.x & 31) < active_lanes) {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = fma(a[i], m, c);
    }
  }
  long long t1 = clock64();
  double s = a[0] + a[1] + a[2] + a[3];
  if (s == 12345.678) out[0] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
  for (int lanes : {32, 16, 8, 24}) {
    chain<<<1, 32 * 4 * 6>>>(4096, lanes, out, cyc);  // 6 warps per SMSP, ILP 4: pipe-bound
    chain<<<1, 32 * 4 * 6>>>(4096, lanes, out, cyc);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("active lanes %2d: %.2f cycles per warp-DFMA per SMSP\n", lanes, (double)h / (4096.0 * 4 * 6));
  }
  return 0;
}
