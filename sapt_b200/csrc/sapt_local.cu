// Per-chain "local" surrogate of the reference's fork localsurrogates/surrogate_pt_classifier_rw_localsurrogate.py
// (LS): a 4 -> 10 -> 1 sigmoid network per chain, trained online on lag-2 windows of the chain's own likelihood
// history and asked for the next tempered likelihood on surrogate steps. All arithmetic in fp64 like the reference.
//
//   ls_train_kernel   one warp per chain; lane h < 10 owns hidden unit h (its W1 column, W2 entry, B1 entry) in
//                     registers; the 200 x windows per-row SGD steps (backprop_train LS:152-177) are sequential,
//                     each one is a 4-term dot + sigmoid per lane, a 16-lane shuffle reduction, a sigmoid, updates.
//   ls_predict_kernel one thread per surrogate row (surrogate.predict LS:313-325).
#include <math.h>

#include "sapt_device.cuh"
#include "sapt_internal.h"

namespace sapt {

namespace {
constexpr int kIn = SAPT_LS_INPUTS, kHid = SAPT_LS_HIDDEN, kW = SAPT_LS_WEIGHTS, kBlock = 5, kLag = 2;
constexpr int kSeriesSmem = 4096;   // series entries staged in shared memory (longer series are read from global)

__device__ __forceinline__ double sigmoid_d(double z) { return 1.0 / (1.0 + exp(-z)); }

__global__ void __launch_bounds__(32) ls_train_kernel(sapt_local_surrogate m, int row_lo, int row_hi, int first, const double* init_inject,
                                                      uint32_t seed, double lrate, int epochs, double* rmse) {
  __shared__ double s_y[kSeriesSmem];
  const int row = blockIdx.x;
  if (row >= m.R) return;
  const int slot = m.slot_of_row[row];
  const int lane = threadIdx.x;
  const int n_y = row_hi - row_lo;
  const int trim = n_y % kBlock;                       // LS:774-776
  const int len = n_y - trim;
  const int nb = (len - kBlock) / kLag;                // LS:781 (host guarantees len >= kBlock)
  const double* col = m.surg + ((int64_t)(row_lo + trim) * m.n_slots + slot) * 3 + 2;
  const int64_t stride = (int64_t)m.n_slots * 3;
  double limit;
  if (first) {                                         // LS:791: limit = 2 * min(y)
    double mn = INFINITY;
    for (int j = lane; j < len; j += 32) mn = fmin(mn, col[j * stride]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    limit = mn * 2.0;
    if (lane == 0) m.limit[slot] = limit;
  } else {
    limit = m.limit[slot];
  }
  const bool staged = len <= kSeriesSmem;
  if (staged) for (int j = lane; j < len; j += 32) s_y[j] = col[j * stride] / limit;   // LS:797 / LS:829
  __syncwarp();
  auto yv = [&](int j) { return staged ? s_y[j] : col[j * stride] / limit; };

  double* w = m.weights + (int64_t)slot * kW;
  const bool act = lane < kHid;
  const int h = act ? lane : 0;
  double w1[kIn], w2, b1, b2;
  if (first) {                                         // create_network LS:62-65: W1, B1, W2, B2 in drawing order
    auto zn = [&](int idx) -> double {
      if (init_inject) return init_inject[(int64_t)slot * kW + idx];
      float z[4];
      philox_normal4(seed, SAPT_STREAM_LS_INIT, 0, (uint32_t)slot, (uint32_t)(idx >> 2), z);
      return (double)z[idx & 3];
    };
#pragma unroll
    for (int k = 0; k < kIn; ++k) w1[k] = zn(k * kHid + h) / sqrt((double)kIn);
    b1 = zn(kIn * kHid + h) / sqrt((double)kHid);
    w2 = zn(kIn * kHid + kHid + h) / sqrt((double)kHid);
    b2 = zn(kIn * kHid + 2 * kHid) / sqrt((double)kHid);
  } else {
#pragma unroll
    for (int k = 0; k < kIn; ++k) w1[k] = w[k * kHid + h];
    w2 = w[kIn * kHid + h];
    b1 = w[kIn * kHid + kHid + h];
    b2 = w[kIn * kHid + 2 * kHid];
  }
  double sse = 0.0;
  for (int e = 0; e < epochs; ++e) {
    for (int b = 0; b < nb; ++b) {
      double x[kIn];
#pragma unroll
      for (int k = 0; k < kIn; ++k) x[k] = yv(b * kLag + k);
      const double target = yv(b * kLag + kIn);
      double z1 = 0.0;
#pragma unroll
      for (int k = 0; k < kIn; ++k) z1 += x[k] * w1[k];
      const double hid = sigmoid_d(z1 - b1);                                  // ForwardPass LS:80-84
      double part = act ? hid * w2 : 0.0;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);   // lanes 0..15 hold the sum
      part = __shfl_sync(0xffffffffu, part, 0);
      const double out = sigmoid_d(part - b2);
      // BackwardPass LS:90-101: desired = one_hot(int(target)) over a single output neuron = 1.0
      const double out_delta = (1.0 - out) * (out * (1.0 - out));
      const double hid_delta = (out_delta * w2) * (hid * (1.0 - hid));       // W2 before its update
      w2 += hid * (out_delta * lrate);
      b2 += -1.0 * lrate * out_delta;
#pragma unroll
      for (int k = 0; k < kIn; ++k) w1[k] += x[k] * (hid_delta * lrate);
      b1 += -1.0 * lrate * hid_delta;
      const double err = out - target;                                        // sampleEr LS:75-78 on the raw target
      sse += err * err;
    }
  }
  if (act) {
#pragma unroll
    for (int k = 0; k < kIn; ++k) w[k * kHid + h] = w1[k];
    w[kIn * kHid + h] = w2;
    w[kIn * kHid + kHid + h] = b1;
  }
  if (lane == 0) {
    w[kIn * kHid + 2 * kHid] = b2;
    if (rmse) rmse[slot] = sqrt(sse / (double)nb);                            // LS:173 (nan when there is no window)
  }
}

__global__ void ls_predict_kernel(sapt_local_surrogate m, int step, const int32_t* step_dev, const int32_t* rows, const int32_t* n_rows,
                                  double* pred, double* avg) {
  const int nact = n_rows ? *n_rows : m.R;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nact) return;
  const int row = rows ? rows[t] : t;
  const int slot = m.slot_of_row[row];
  const int i = step + (step_dev ? *step_dev : 0);
  const double limit = m.limit[slot];
  const double* w = m.weights + (int64_t)slot * kW;
  double x[kIn];
#pragma unroll
  for (int k = 0; k < kIn; ++k) x[k] = m.surg[((int64_t)(i - kIn + k) * m.n_slots + slot) * 3 + 2] / limit;   // LS:592
  double z2 = 0.0;
  for (int h = 0; h < kHid; ++h) {
    double z1 = 0.0;
#pragma unroll
    for (int k = 0; k < kIn; ++k) z1 += x[k] * w[k * kHid + h];
    z2 += sigmoid_d(z1 - w[kIn * kHid + kHid + h]) * w[kIn * kHid + h];
  }
  const double out = sigmoid_d(z2 - w[kIn * kHid + 2 * kHid]);
  pred[row] = out * limit;                                                    // LS:598
  avg[row] = ((((x[0] + x[1]) + x[2]) + x[3]) / (double)kIn) * limit;        // np.mean(X_load) * limit, LS:617
}
}  // namespace

}  // namespace sapt

using namespace sapt;

extern "C" int sapt_local_surrogate_train(const sapt_local_surrogate* m, int row_lo, int row_hi, int first, const double* init_inject,
                                          uint32_t seed, double lrate, int epochs, double* rmse, void* stream) {
  SAPT_REQUIRE(m && m->weights && m->limit && m->surg && m->slot_of_row && m->n_slots > 0 && m->R >= 0, "sapt_local_surrogate_train");
  SAPT_REQUIRE(row_lo >= 0 && row_hi - row_lo - (row_hi - row_lo) % 5 >= 5, "local surrogate: the series needs at least five entries (LS:781)");
  SAPT_REQUIRE(epochs >= 0, "epochs");
  if (m->R == 0) return SAPT_OK;
  ls_train_kernel<<<m->R, 32, 0, (cudaStream_t)stream>>>(*m, row_lo, row_hi, first, init_inject, seed, lrate, epochs, rmse);
  SAPT_CHECK_LAUNCH();
  return SAPT_OK;
}

extern "C" int sapt_local_surrogate_predict(const sapt_local_surrogate* m, int step, const int32_t* step_dev, const int32_t* rows,
                                            const int32_t* n_rows, double* pred, double* avg, void* stream) {
  SAPT_REQUIRE(m && m->weights && m->limit && m->surg && m->slot_of_row && pred && avg, "sapt_local_surrogate_predict");
  SAPT_REQUIRE((rows == nullptr) == (n_rows == nullptr), "rows and n_rows come together");
  SAPT_REQUIRE(step_dev || step >= SAPT_LS_INPUTS, "local surrogate: predict needs four history rows");
  if (m->R == 0) return SAPT_OK;
  ls_predict_kernel<<<(m->R + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*m, step, step_dev, rows, n_rows, pred, avg);
  SAPT_CHECK_LAUNCH();
  return SAPT_OK;
}
