// Internal declarations shared by the translation units of libsapt.so (not part of the ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "sapt.h"

namespace sapt {

// ---- error plumbing -------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define SAPT_CUDA(call)                                                                   \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      ::sapt::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return SAPT_ERR_CUDA;                                                               \
    }                                                                                     \
  } while (0)

#define SAPT_CHECK_LAUNCH()                                                               \
  do {                                                                                    \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess) {                                                             \
      ::sapt::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return SAPT_ERR_CUDA;                                                               \
    }                                                                                     \
    ::sapt::count_launch();                                                               \
  } while (0)

#define SAPT_REQUIRE(cond, msg)                                                           \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      ::sapt::set_error("invalid argument: %s (%s)", msg, #cond);                         \
      return SAPT_ERR_INVALID;                                                            \
    }                                                                                     \
  } while (0)

// ---- data set resident in HBM ---------------------------------------------------------
struct DataSet {
  int64_t n = 0;        // rows
  float* x = nullptr;   // [n_pad][DP] fp32, zero padded (features beyond D and rows beyond n)
  int32_t* y = nullptr; // [n_pad] labels; padding rows carry -1
  int64_t n_pad = 0;    // n rounded up to a multiple of 256 (tile) so kernels never branch on loads
  // tcgen05 operand image: [n_pad][KA] bf16, KA = 64: [x_hi | x_lo | x_hi | 1 1 0..] (see sapt_loglik_tc.cu)
  void* a_bf16 = nullptr;
  void* tmap = nullptr; // CUtensorMap (128 B, 64-byte aligned host copy) for a_bf16
  // likelihood kernel's operand image: [n_pad][64 KBA] fp16 = [x_hi | x_lo | (1 1 0..)] (two fp16 pieces, see sapt_loglik_tc2.cu)
  void* a_f16 = nullptr;
  void* tmap_f16 = nullptr;
  bool f16_range_ok = true;  // every |feature| <= 6e4: representable by the fp16 pieces (else the tcgen05 likelihood path is refused for this set)
};

}  // namespace sapt

struct sapt_problem {
  int D, H, O, P;
  int DP;   // D padded: 8,16,32,64
  int HP;   // H padded to a multiple of 4
  int OP;   // O padded to a multiple of 4
  int device;
  int num_sms;
  sapt::DataSet set[2];
  // workspace for per-(row-chunk, replica) partial sums, grown on demand
  void* ws = nullptr;
  size_t ws_bytes = 0;
  // pinned + device staging for the *_host entry points
  float* h2d_theta = nullptr; size_t h2d_theta_bytes = 0;
  double* d_L = nullptr; int32_t* d_nc = nullptr; double* d_sq = nullptr; int out_cap = 0;
  bool tc_ok = false;
};

namespace sapt {
int ensure_workspace(sapt_problem* p, size_t bytes);
// implemented in sapt_loglik_f32.cu
int launch_loglik_f32(sapt_problem* p, int which, const float* theta, int64_t ld, int R,
                      const int32_t* rows, const int32_t* n_rows, double* L, int32_t* nc, double* sq,
                      cudaStream_t st);
int launch_loglik_finalize(const double* partL, const int32_t* partNc, const int32_t* partSq, int n_chunks, int R, const int32_t* rows,
                           const int32_t* n_rows, double* L, int32_t* nc, double* sq, cudaStream_t st);
int launch_forward_rows_f32(sapt_problem* p, int which, const float* theta, int32_t* fx, float* prob, cudaStream_t st);
// implemented in sapt_loglik_tc.cu
int tc_prepare_dataset(sapt_problem* p, int which, const float* x_host, int64_t n);
int launch_loglik_tc(sapt_problem* p, int which, const float* theta, int64_t ld, int R,
                     const int32_t* rows, const int32_t* n_rows, double* L, int32_t* nc, double* sq,
                     cudaStream_t st);
bool tc_topology_supported(int D, int H, int O);
}  // namespace sapt
