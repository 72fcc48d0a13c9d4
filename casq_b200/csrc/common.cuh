// casq_b200 internal: shared declarations for the host model and the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "../../include/casq_b200.h"

typedef std::complex<double> cplx;

void casq_set_error(const char* fmt, ...);

#define CASQ_FAIL(code, ...)     \
  do {                           \
    casq_set_error(__VA_ARGS__); \
    return (code);               \
  } while (0)

#define CASQ_CUDA(expr)                                                                        \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      casq_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return CASQ_ERR_CUDA;                                                                    \
    }                                                                                          \
  } while (0)

// Time grid of one fixed-step solve, shared by every point of the batch (upstream
// get_fixed_step_sizes + the `t += h` accumulation of fixed_step_solver_template).
struct TimeGrid {
  std::vector<int> piece_nsteps;    // per piece
  std::vector<double> piece_h;      // per piece
  std::vector<double> out_times;    // merged t args
  std::vector<char> out_keep;       // which merged times are returned (t_eval trimming)
  std::vector<double> times;        // per piece: 2n+1 stage times (t, t+h/2, t+h, ...)
  std::vector<int> sample_idx;      // floor_divide(t, dt) clipped to [-1, n_total]
};

int casq_build_time_grid(double t0, double t1, double max_dt, int T, const double* t_eval, double dt,
                         int n_total_samples, TimeGrid* g);
long long casq_py_floor_divide(double a, double b);

struct DeviceBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int ensure(size_t need);
  void release();
};

struct casq_model {
  int d = 0, K = 0, device = 0;
  double dt = 0;
  bool rwa = false;
  double rwa_cutoff = 0;
  int frame_kind = 0;
  std::vector<cplx> H0, Hk;  // lab
  std::vector<double> nu;    // carrier Hz
  std::vector<double> fe;    // frame eigenvalues [d]
  std::vector<cplx> U;       // frame basis [d*d], columns = eigenvectors
  bool U_is_identity = true;
  std::vector<cplx> S;       // frame-basis static (H0 - Hf), RWA-masked   [d*d]
  std::vector<cplx> P, N;    // per channel: multiplies z / conj(z)        [K*d*d]; no-RWA: both = G/2
  std::vector<double> de;    // dressed evals
  std::vector<cplx> V;       // dressed states
  int J = 0;
  std::vector<cplx> Lfb;     // dissipators in frame basis [J*d*d]
  // device scratch (lazily grown)
  DeviceBuf tab, tab_idx, tab_times, pieces_n, pieces_h, consts, work0, work1, work2, work3;
  // cache key of the table currently resident in `tab`
  std::string tab_key;
  long long tab_NT = 0;   // shape of the cached fixed-step grid (small RK4 path)
  int tab_n_pieces = 0, tab_piece0_steps = 0;
  DeviceBuf slice_buf;    // ticket queue and hand-over state of the time-sliced headline kernel
  // constant-table variant of the headline kernel (sv_small.cu, ConstTab): offset table + frame phasors (host, cached with
  // the grid) and the per-chunk phasor table (device)
  std::vector<double> small_ct;
  DeviceBuf ct_chunk;
  // pinned host staging of the small RK4 path's grid upload (pageable cudaMemcpyAsync synchronises the stream per copy)
  void* stage_h = nullptr;
  size_t stage_bytes = 0;
  cudaEvent_t stage_ev = nullptr;  // recorded after the last copy that reads stage_h
  // second stream + fork/join events of the Lindblad path (several parts of the batch in flight fill each other's grid tails)
  cudaStream_t aux_stream[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
};

// hermitian eigen-decomposition (cyclic complex Jacobi); eigenvalues ascending, columns of V.
void casq_eigh(int n, const cplx* A, double* evals, cplx* V);
