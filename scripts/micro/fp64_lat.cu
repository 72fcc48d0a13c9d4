// microbenchmark: DFMA dependent latency and per-SMSP throughput vs (warps, ILP) on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void chain(int iters, double* out, long long* cyc) {
  double a[ILP];
  for (int i = 0; i < ILP; i++) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
  const double m = 1.0 - 1e-12, c = 1e-13;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = fma(a[i], m, c);
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < ILP; i++) s += a[i];
  if (s == 12345.678) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP>
void run(int warps_per_smsp) {
  double* out; long long* cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
  int iters = 4096;
  int threads = 32 * 4 * warps_per_smsp;  // one block on one SM
  chain<ILP><<<1, threads>>>(iters, out, cyc);
  chain<ILP><<<1, threads>>>(iters, out, cyc);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double per_instr_per_smsp = (double)h / ((double)iters * ILP * warps_per_smsp);
  printf("ILP=%d warps/SMSP=%d: %.2f cycles per iter, %.2f cycles per warp-DFMA per SMSP\n", ILP, warps_per_smsp, (double)h / iters, per_instr_per_smsp);
}
int main() {
  run<1>(1); run<2>(1); run<4>(1); run<8>(1);
  run<1>(2); run<1>(4); run<1>(5); run<1>(6); run<1>(8);
  run<2>(5); run<3>(5); run<4>(5); run<2>(6); run<4>(6);
  return 0;
}
