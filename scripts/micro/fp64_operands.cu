// fp64 pipe rate with REALISTIC operands: every DFMA/DMUL reads three (two) distinct, changing registers, as the RK4 kernels
// do, against the same instruction count with loop-invariant multiplicands (register-reuse friendly).  12 independent
// accumulator chains per thread (the headline kernel has 2 trajectories x 6 components), 1..4 warps per scheduler.
// Build: nvcc -cudart shared -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_microbin/fp64_operands scripts/micro/fp64_operands.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(64) k(double* out, const double* in, int iters) {
  double acc[12], x[8], y[8];
  for (int i = 0; i < 12; i++) acc[i] = in[threadIdx.x + 64 * i];
  for (int i = 0; i < 8; i++) {
    x[i] = in[threadIdx.x + 64 * (12 + i)];
    y[i] = in[threadIdx.x + 64 * (20 + i)];
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < 12; i++) {
        if (MODE == 0) acc[i] = fma(x[0], y[0], acc[i]);                           // invariant multiplicands
        if (MODE == 1) acc[i] = fma(x[(i + r) & 7], y[(3 * i + r) & 7], acc[i]);   // distinct operands
        if (MODE == 2) acc[i] = fma(x[(i + r) & 7], acc[(i + 5) % 12], acc[i]);    // operands produced by other chains (matvec-like)
      }
      if (MODE == 2) {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = x[i] * 1.0000001;  // ~1 DMUL per 6 DFMA keeps x changing
      }
    }
  }
  double s = 0;
  for (int i = 0; i < 12; i++) s += acc[i];
  for (int i = 0; i < 8; i++) s += x[i];
  out[blockIdx.x * 64 + threadIdx.x] = s;
}

// MODE 3: the multiplicand comes from a table passed as a kernel PARAMETER (constant bank -> LDCU -> uniform register
// operand of the DFMA), the other two operands are distinct vector registers
struct ParamTab { double t[512]; };
__global__ void __launch_bounds__(64) kparam(double* out, const double* in, const ParamTab p, int iters) {
  double acc[12], x[8];
  for (int i = 0; i < 12; i++) acc[i] = in[threadIdx.x + 64 * i];
  for (int i = 0; i < 8; i++) x[i] = in[threadIdx.x + 64 * (12 + i)];
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const double c = p.t[(it * 4 + r) & 511];
#pragma unroll
      for (int i = 0; i < 12; i++) acc[i] = fma(c, x[(i + r) & 7], acc[i]);
    }
  }
  double s = 0;
  for (int i = 0; i < 12; i++) s += acc[i];
  out[blockIdx.x * 64 + threadIdx.x] = s;
}
void run_param(int blocks_per_sm, double* out, double* in) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 20000;
  ParamTab p;
  for (int i = 0; i < 512; i++) p.t[i] = 1e-9 * i;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  kparam<<<sms * blocks_per_sm, 64>>>(out, in, p, 100);
  cudaEventRecord(e0);
  kparam<<<sms * blocks_per_sm, 64>>>(out, in, p, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double warps_per_smsp = blocks_per_sm * 2 / 4.0;
  printf("%-44s warps/SMSP %.1f: %.3f ms, %.2f cycles per fp64 instruction per SMSP\n", "multiplicand from a kernel parameter (UR)", warps_per_smsp, ms,
         ms * 1e-3 * clk * 1e3 / ((double)iters * 48 * warps_per_smsp));
}

template <int MODE>
void run(const char* name, int blocks_per_sm, double* out, double* in) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<sms * blocks_per_sm, 64>>>(out, in, 100);
  cudaEventRecord(e0);
  k<MODE><<<sms * blocks_per_sm, 64>>>(out, in, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double fp64_per_thread = (double)iters * 4 * (12 + (MODE == 2 ? 8 : 0));
  const double warps_per_smsp = blocks_per_sm * 2 / 4.0;
  const double cycles = ms * 1e-3 * clk * 1e3;
  printf("%-44s warps/SMSP %.1f: %.3f ms, %.2f cycles per fp64 instruction per SMSP\n", name, warps_per_smsp, ms,
         cycles / (fp64_per_thread * warps_per_smsp));
}

int main() {
  double *out, *in;
  cudaMalloc(&out, 148 * 16 * 64 * 8);
  cudaMalloc(&in, 64 * 32 * 8);
  cudaMemset(in, 0, 64 * 32 * 8);
  for (int b : {2, 4, 6, 8}) {
    run<0>("invariant multiplicands", b, out, in);
    run<1>("distinct register operands", b, out, in);
    run<2>("operands from other chains + DMUL stream", b, out, in);
    run_param(b, out, in);
  }
  return 0;
}
