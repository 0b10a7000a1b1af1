// Langevin "gradient" of the reference = one epoch of per-row online SGD
// (misc_codeapril2020/surrogate_pt_classifier_langevingrad.py:138-156 + BackwardPass :96-106).
//
// mode 0 "sgd_exact": reference semantics. The recurrence is sequential in the rows (row i sees
//   the weights row i-1 left behind), so the parallelism is chains x neurons: one CTA per chain,
//   weights resident in shared memory in fp64 (the loop is latency bound, fp64 costs nothing here
//   and keeps the result within 1e-10 of the numpy epoch), four barriers per row.
// mode 1 "batch": theta + lr*scale * sum_i g_i(theta) at FIXED theta — the transposed contraction
//   X^T.delta1 / hid^T.delta2. grid = (row chunks, chains); per 128-row tile the CTA
//   computes activations/deltas one row per thread, then accumulates the outer products with the
//   tile staged in shared memory; chunk partials are folded by a second kernel.
#include <math.h>

#include <algorithm>

#include "sapt_device.cuh"
#include "sapt_internal.h"

namespace sapt {

__global__ void __launch_bounds__(256) langevin_sgd_kernel(const float* x, const int32_t* y, int64_t n, int DP, int D, int H, int O,
                                                           const float* theta, float* theta_gd, int64_t ld, const int32_t* rows,
                                                           const int32_t* n_rows, int R, double lr) {
  const int nact = n_rows ? *n_rows : R;
  if ((int)blockIdx.x >= nact) return;
  const int row = rows ? rows[blockIdx.x] : (int)blockIdx.x;
  extern __shared__ double sm[];
  const int P = D * H + H * O + H + O;
  double* W1 = sm;             // [D][H]
  double* W2 = W1 + D * H;     // [H][O]
  double* B1 = W2 + H * O;     // [H]
  double* B2 = B1 + H;         // [O]
  double* xs = B2 + O;         // [D]
  double* hid = xs + D;        // [H]
  double* hdel = hid + H;      // [H]
  double* odel = hdel + H;     // [O]
  const int tid = threadIdx.x, nt = blockDim.x;
  const float* src = theta + (int64_t)row * ld;
  for (int i = tid; i < P; i += nt) sm[i] = (double)src[i];
  __syncthreads();
  for (int64_t r = 0; r < n; ++r) {
    for (int d = tid; d < D; d += nt) xs[d] = (double)x[r * DP + d];
    const int label = y[r];
    __syncthreads();
    for (int h = tid; h < H; h += nt) {                       // ForwardPass, LG:70-72
      double z = 0.0;
      for (int d = 0; d < D; ++d) z += xs[d] * W1[d * H + h];
      z -= B1[h];
      hid[h] = 1.0 / (1.0 + exp(-z));
    }
    __syncthreads();
    for (int o = tid; o < O; o += nt) {                       // LG:73-74, LG:100
      double z = 0.0;
      for (int h = 0; h < H; ++h) z += hid[h] * W2[h * O + o];
      z -= B2[o];
      const double out = 1.0 / (1.0 + exp(-z));
      const double desired = (o == label) ? 1.0 : 0.0;
      odel[o] = (desired - out) * (out * (1.0 - out));
    }
    __syncthreads();
    for (int h = tid; h < H; h += nt) {                       // LG:101 (W2 BEFORE its update)
      double s = 0.0;
      for (int o = 0; o < O; ++o) s += odel[o] * W2[h * O + o];
      hdel[h] = s * (hid[h] * (1.0 - hid[h]));
    }
    __syncthreads();
    for (int i = tid; i < H * O; i += nt) { const int h = i / O, o = i - h * O; W2[i] += hid[h] * (odel[o] * lr); }   // LG:102
    for (int o = tid; o < O; o += nt) B2[o] += (-1.0 * lr) * odel[o];                                                 // LG:103
    for (int i = tid; i < D * H; i += nt) { const int d = i / H, h = i - d * H; W1[i] += xs[d] * (hdel[h] * lr); }    // LG:105
    for (int h = tid; h < H; h += nt) B1[h] += (-1.0 * lr) * hdel[h];                                                 // LG:106
    __syncthreads();
  }
  float* dst = theta_gd + (int64_t)row * ld;
  for (int i = tid; i < P; i += nt) dst[i] = (float)sm[i];
}

constexpr int kBT = 128;  // rows per tile == threads

__global__ void __launch_bounds__(kBT) langevin_batch_kernel(const float* x, const int32_t* y, int64_t n, int DP, int D, int H, int O,
                                                             const float* theta, int64_t ld, const int32_t* rows, const int32_t* n_rows,
                                                             int R, int tiles_per_chunk, float* part /*[chunks][R][P]*/) {
  const int nact = n_rows ? *n_rows : R;
  if ((int)blockIdx.y >= nact) return;
  const int row = rows ? rows[blockIdx.y] : (int)blockIdx.y;
  extern __shared__ float smf[];
  const int P = D * H + H * O + H + O;
  const int XS = D | 1, HS = H | 1, OS = O | 1;  // odd strides: conflict-free row-per-thread access
  float* w = smf;                 // [P] theta (W1,W2,B1,B2)
  float* g = w + P;               // [P] gradient accumulator
  float* xt = g + P;              // [kBT][XS]
  float* ht = xt + kBT * XS;      // [kBT][HS] hidden activations
  float* hd = ht + kBT * HS;      // [kBT][HS] hidden deltas
  float* od = hd + kBT * HS;      // [kBT][OS] output deltas
  const float* W1 = w; const float* W2 = w + D * H; const float* B1 = W2 + H * O; const float* B2 = B1 + H;
  const int tid = threadIdx.x;
  const float* src = theta + (int64_t)row * ld;
  for (int i = tid; i < P; i += kBT) { w[i] = src[i]; g[i] = 0.f; }
  __syncthreads();
  const int64_t n_tiles = (n + kBT - 1) / kBT;
  const int64_t t0 = (int64_t)blockIdx.x * tiles_per_chunk, t1 = min(t0 + tiles_per_chunk, n_tiles);
  for (int64_t t = t0; t < t1; ++t) {
    const int64_t r = t * kBT + tid;
    const bool valid = r < n;
    // phase A: one row per thread
    for (int d = 0; d < D; ++d) xt[tid * XS + d] = valid ? x[r * DP + d] : 0.f;
    const int label = valid ? y[r] : -1;
    for (int h = 0; h < H; ++h) {
      float z = -B1[h];
      for (int d = 0; d < D; ++d) z = fmaf(xt[tid * XS + d], W1[d * H + h], z);
      ht[tid * HS + h] = 1.0f / (1.0f + __expf(-z));
    }
    for (int o = 0; o < O; ++o) {
      float z = -B2[o];
      for (int h = 0; h < H; ++h) z = fmaf(ht[tid * HS + h], W2[h * O + o], z);
      const float out = 1.0f / (1.0f + __expf(-z));
      const float desired = (o == label) ? 1.f : 0.f;
      od[tid * OS + o] = valid ? (desired - out) * (out * (1.f - out)) : 0.f;
    }
    for (int h = 0; h < H; ++h) {
      float s = 0.f;
      for (int o = 0; o < O; ++o) s = fmaf(od[tid * OS + o], W2[h * O + o], s);
      const float hv = ht[tid * HS + h];
      hd[tid * HS + h] = valid ? s * (hv * (1.f - hv)) : 0.f;
    }
    __syncthreads();
    // phase B: outer products reduced over the tile's rows
    for (int e = tid; e < P; e += kBT) {
      float acc = 0.f;
      if (e < D * H) {
        const int d = e / H, h = e - d * H;
        for (int rr = 0; rr < kBT; ++rr) acc = fmaf(xt[rr * XS + d], hd[rr * HS + h], acc);
      } else if (e < D * H + H * O) {
        const int j = e - D * H; const int h = j / O, o = j - h * O;
        for (int rr = 0; rr < kBT; ++rr) acc = fmaf(ht[rr * HS + h], od[rr * OS + o], acc);
      } else if (e < D * H + H * O + H) {
        const int h = e - D * H - H * O;
        for (int rr = 0; rr < kBT; ++rr) acc -= hd[rr * HS + h];
      } else {
        const int o = e - D * H - H * O - H;
        for (int rr = 0; rr < kBT; ++rr) acc -= od[rr * OS + o];
      }
      g[e] += acc;
    }
    __syncthreads();
  }
  float* dst = part + ((int64_t)blockIdx.x * R + blockIdx.y) * P;
  for (int i = tid; i < P; i += kBT) dst[i] = g[i];
}

__global__ void langevin_batch_finalize(const float* part, int n_chunks, int R, int P, const float* theta, float* theta_gd, int64_t ld,
                                        const int32_t* rows, const int32_t* n_rows, double lr_scale) {
  const int nact = n_rows ? *n_rows : R;
  if ((int)blockIdx.y >= nact) return;
  const int row = rows ? rows[blockIdx.y] : (int)blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  double s = 0.0;
  for (int c = 0; c < n_chunks; ++c) s += (double)part[((int64_t)c * R + blockIdx.y) * P + i];
  theta_gd[(int64_t)row * ld + i] = (float)((double)theta[(int64_t)row * ld + i] + lr_scale * s);
}

}  // namespace sapt

using namespace sapt;

extern "C" int sapt_langevin(sapt_problem_t* p, const float* theta, float* theta_gd, int64_t ld, int R, const int32_t* rows,
                             const int32_t* n_rows, double lrate, int mode, double batch_scale, void* stream) {
  SAPT_REQUIRE(p && theta && theta_gd && R >= 0 && ld >= p->P, "sapt_langevin");
  if (R == 0) return SAPT_OK;
  const DataSet& ds = p->set[0];
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 0) {
    const size_t smem = (size_t)(p->P + p->D + 2 * p->H + p->O) * sizeof(double);
    if (smem > 200 * 1024) { set_error("sapt_langevin: topology too large for the sgd_exact kernel"); return SAPT_ERR_UNSUPPORTED; }
    if (smem > 48 * 1024) SAPT_CUDA(cudaFuncSetAttribute(langevin_sgd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    langevin_sgd_kernel<<<R, 256, smem, st>>>(ds.x, ds.y, ds.n, p->DP, p->D, p->H, p->O, theta, theta_gd, ld, rows, n_rows, R, lrate);
    SAPT_CHECK_LAUNCH();
    return SAPT_OK;
  }
  if (mode == 1) {
    const int P = p->P;
    const size_t smem = (size_t)(2 * P + kBT * ((p->D | 1) + 2 * (p->H | 1) + (p->O | 1))) * sizeof(float);
    if (smem > 220 * 1024) { set_error("sapt_langevin: topology too large for the batch kernel"); return SAPT_ERR_UNSUPPORTED; }
    const int64_t n_tiles = (ds.n + kBT - 1) / kBT;
    int64_t n_chunks = std::min<int64_t>(n_tiles, std::max<int64_t>(1, ((int64_t)p->num_sms * 4 + R - 1) / R));
    const int tiles_per_chunk = (int)((n_tiles + n_chunks - 1) / n_chunks);
    n_chunks = (n_tiles + tiles_per_chunk - 1) / tiles_per_chunk;
    int rc = ensure_workspace(p, (size_t)n_chunks * R * P * sizeof(float));
    if (rc) return rc;
    if (smem > 48 * 1024) SAPT_CUDA(cudaFuncSetAttribute(langevin_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    langevin_batch_kernel<<<dim3((unsigned)n_chunks, (unsigned)R), kBT, smem, st>>>(ds.x, ds.y, ds.n, p->DP, p->D, p->H, p->O, theta, ld, rows,
                                                                                      n_rows, R, tiles_per_chunk, (float*)p->ws);
    SAPT_CHECK_LAUNCH();
    langevin_batch_finalize<<<dim3((P + 127) / 128, R), 128, 0, st>>>((const float*)p->ws, (int)n_chunks, R, P, theta, theta_gd, ld, rows, n_rows,
                                                                      lrate * batch_scale);
    SAPT_CHECK_LAUNCH();
    return SAPT_OK;
  }
  set_error("sapt_langevin: unknown mode %d", mode);
  return SAPT_ERR_INVALID;
}
