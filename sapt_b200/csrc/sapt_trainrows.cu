// Surrogate rendezvous (RW:621-657, RW:1105-1116; LG:819-844): moving the interval's (theta', L) training rows.
// The reference pickles an (n, P+1) float64 block per chain through a multiprocessing queue and concatenates them
// chain by chain. Here the rows sit in slot-indexed device buffers on whichever rank hosted the chain when they were
// written; two kernels do all row movement, driven by index vectors the host logic (parallel.py) computes with
// static shapes — no data-dependent allocation, no host synchronisation:
//
//   train_rows_pack    payload[i] = [ theta' (P fp32) | L (fp64 as two words) | destination row | 0 ]  for the i-th
//                      row this rank owns (src[i] = slot * cap + j), or destination -1 for the padding rows
//   train_rows_scatter batch[dest] = (double) payload row, for the payloads of all ranks (gathered on the trainer
//                      rank), plus the all-zero first row of every chain the RW / LG forks ship (RW:496,626)
//
// One CTA per row; the copies are coalesced over P. HBM-bound: 4 (P + 4) bytes read + written per row in pack,
// 4 (P + 4) read + 8 (P + 1) written in scatter.
#include "sapt_device.cuh"
#include "sapt_internal.h"

namespace sapt {

__global__ void train_rows_pack_kernel(const float* __restrict__ train_theta, const double* __restrict__ train_L, const int32_t* __restrict__ src,
                                       const int32_t* __restrict__ dest, int P, float* __restrict__ payload) {
  const int64_t i = blockIdx.x;
  const int W = P + 4;
  float* out = payload + i * W;
  const int s = src[i];
  if (s < 0) {                      // padding row: only the destination word is read downstream
    if (threadIdx.x == 0) reinterpret_cast<int32_t*>(out)[P + 2] = -1;
    return;
  }
  const float* in = train_theta + (int64_t)s * P;
  for (int c = threadIdx.x; c < P; c += blockDim.x) out[c] = in[c];
  if (threadIdx.x == 0) {
    const long long bits = __double_as_longlong(train_L[s]);
    reinterpret_cast<int32_t*>(out)[P] = (int32_t)(bits & 0xffffffffll);
    reinterpret_cast<int32_t*>(out)[P + 1] = (int32_t)(bits >> 32);
    reinterpret_cast<int32_t*>(out)[P + 2] = dest[i];
    out[P + 3] = 0.f;
  }
}

__global__ void train_rows_scatter_kernel(const float* __restrict__ payload, int64_t n_payload, int P, const int64_t* __restrict__ zero_rows,
                                          int n_zero, double* __restrict__ batch) {
  const int64_t i = blockIdx.x;
  const int W = P + 4;
  if (i >= n_payload) {             // the all-zero first row of chain i - n_payload
    double* out = batch + zero_rows[i - n_payload] * (P + 1);
    for (int c = threadIdx.x; c <= P; c += blockDim.x) out[c] = 0.0;
    return;
  }
  const float* in = payload + i * W;
  const int d = reinterpret_cast<const int32_t*>(in)[P + 2];
  if (d < 0) return;
  double* out = batch + (int64_t)d * (P + 1);
  for (int c = threadIdx.x; c < P; c += blockDim.x) out[c] = (double)in[c];
  if (threadIdx.x == 0) {
    const unsigned long long lo = (uint32_t)reinterpret_cast<const int32_t*>(in)[P];
    const unsigned long long hi = (uint32_t)reinterpret_cast<const int32_t*>(in)[P + 1];
    out[P] = __longlong_as_double((long long)(lo | (hi << 32)));
  }
}

// The index vectors of the rendezvous in ONE launch (parallel.training_row_plan is the torch restatement the CPU tests run; on the
// device it was two dozen small kernels, 0.4 ms per rendezvous at 8192 chains). One CTA: every thread owns a contiguous range of
// chains, counts the batch rows and the owned rows of its range, a block scan turns the counts into offsets, a second walk writes
//   src[pos] = slot * cap + j, dest[pos] = batch row      for every owned valid entry (pos = 0, 1, ... in chain-major order),
//   zero_rows[slot] = first batch row of the chain        (the all-zero row RW:496,626 ship), n_out = rows of the batch,
//   n_owned = entries this rank owns; src / dest are -1 from n_owned on.
constexpr int kPlanThreads = 1024;
__global__ void __launch_bounds__(kPlanThreads) train_rows_plan_kernel(const int32_t* __restrict__ count, const uint8_t* __restrict__ mine, int Rg, int cap,
                                                                      int n_pad, int lo, int zero, int32_t* __restrict__ src, int32_t* __restrict__ dest,
                                                                      int64_t* __restrict__ zero_rows, int64_t* __restrict__ n_out, int64_t* __restrict__ n_owned) {
  __shared__ long long s_rows[kPlanThreads], s_own[kPlanThreads];
  const int tid = threadIdx.x;
  const int per = (Rg + kPlanThreads - 1) / kPlanThreads;
  const int s0 = min(Rg, tid * per), s1 = min(Rg, s0 + per);
  for (int i = tid; i < n_pad; i += kPlanThreads) { src[i] = -1; dest[i] = -1; }
  long long rows = 0, own = 0;
  for (int s = s0; s < s1; ++s) {
    const int cnt = min(max(count[s], 0), cap);
    rows += max(cnt - lo, 0) + zero;
    for (int j = lo; j < cnt; ++j) own += mine[(int64_t)s * cap + j] ? 1 : 0;
  }
  s_rows[tid] = rows; s_own[tid] = own;
  __syncthreads();
  for (int off = 1; off < kPlanThreads; off <<= 1) {      // inclusive Hillis-Steele scan of both sums
    const long long a = (tid >= off) ? s_rows[tid - off] : 0, b = (tid >= off) ? s_own[tid - off] : 0;
    __syncthreads();
    s_rows[tid] += a; s_own[tid] += b;
    __syncthreads();
  }
  long long start = s_rows[tid] - rows, pos = s_own[tid] - own;   // exclusive offsets of this thread's range
  if (tid == kPlanThreads - 1) { *n_out = s_rows[tid]; *n_owned = s_own[tid]; }
  for (int s = s0; s < s1; ++s) {
    const int cnt = min(max(count[s], 0), cap);
    if (zero_rows) zero_rows[s] = start;
    for (int j = lo; j < cnt; ++j) {
      if (mine[(int64_t)s * cap + j] && pos < n_pad) {     // more owned rows than n_pad cannot happen (see engine); never write out of range
        src[pos] = s * cap + j;
        dest[pos] = (int32_t)(start + zero + (j - lo));
        ++pos;
      }
    }
    start += max(cnt - lo, 0) + zero;
  }
}

}  // namespace sapt

using namespace sapt;

extern "C" int sapt_train_rows_pack(const float* train_theta, const double* train_L, const int32_t* src, const int32_t* dest, int n_rows, int P,
                                    float* payload, void* stream) {
  SAPT_REQUIRE(train_theta && train_L && src && dest && payload && n_rows >= 0 && P > 0, "sapt_train_rows_pack");
  if (n_rows == 0) return SAPT_OK;
  train_rows_pack_kernel<<<n_rows, 128, 0, (cudaStream_t)stream>>>(train_theta, train_L, src, dest, P, payload);
  SAPT_CHECK_LAUNCH();
  return SAPT_OK;
}

extern "C" int sapt_train_rows_scatter(const float* payload, int64_t n_payload, int P, const int64_t* zero_rows, int n_zero, double* batch,
                                       void* stream) {
  SAPT_REQUIRE((payload || n_payload == 0) && batch && n_payload >= 0 && P > 0 && n_zero >= 0 && (n_zero == 0 || zero_rows), "sapt_train_rows_scatter");   // no payload rows at all (an interval of surrogate steps only, several ranks): only the zero rows are written
  const int64_t grid = n_payload + n_zero;
  if (grid == 0) return SAPT_OK;
  SAPT_REQUIRE(grid < (int64_t)1 << 31, "too many rows for one launch");
  train_rows_scatter_kernel<<<(unsigned)grid, 128, 0, (cudaStream_t)stream>>>(payload, n_payload, P, zero_rows, n_zero, batch);
  SAPT_CHECK_LAUNCH();
  return SAPT_OK;
}

extern "C" int sapt_train_rows_plan(const int32_t* train_count, const uint8_t* train_mine, int R_global, int cap, int n_pad, int skip_first, int zero_row,
                                    int32_t* src, int32_t* dest, int64_t* zero_rows, int64_t* n_out, int64_t* n_owned, void* stream) {
  SAPT_REQUIRE(train_count && train_mine && src && dest && n_out && n_owned && R_global >= 0 && cap >= 1 && n_pad >= 0, "sapt_train_rows_plan");
  SAPT_REQUIRE(!zero_row || zero_rows, "sapt_train_rows_plan: zero_rows");
  SAPT_REQUIRE((int64_t)R_global * cap < ((int64_t)1 << 31), "sapt_train_rows_plan: slot * cap + j must fit 32 bits");
  train_rows_plan_kernel<<<1, kPlanThreads, 0, (cudaStream_t)stream>>>(train_count, train_mine, R_global, cap, n_pad, skip_first ? 1 : 0, zero_row ? 1 : 0, src, dest,
                                                                       zero_row ? zero_rows : nullptr, n_out, n_owned);
  SAPT_CHECK_LAUNCH();
  return SAPT_OK;
}
