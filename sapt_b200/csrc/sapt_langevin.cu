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

// ---- sgd_exact, register-resident version (H <= 128): thread h owns column h of W1, row h of W2 and B1[h] in fp64
// registers; every thread keeps its own (identical) copy of B2. One row costs D + O FMAs forward, one O-wide reduction
// over the hidden units (warp shuffles + one shared-memory hop, ONE barrier per row) and D + O FMAs of update, instead
// of four barriers and shared-memory weights. Same arithmetic as the reference epoch (LG:138-156, 96-106).
template <int DP, int OP>
__global__ void __launch_bounds__(128) langevin_sgd_reg_kernel(const float* x, int xstride, const int32_t* y, int64_t n, int D, int H, int O,
                                                               const float* theta, float* theta_gd, int64_t ld, const int32_t* rows,
                                                               const int32_t* n_rows, int R, double lr) {
  const int nact = n_rows ? *n_rows : R;
  if ((int)blockIdx.x >= nact) return;
  const int row = rows ? rows[blockIdx.x] : (int)blockIdx.x;
  __shared__ double xs[3][DP];
  __shared__ int labs[3];
  __shared__ double red[2][4][OP];
  const int h = threadIdx.x, lane = h & 31, warp = h >> 5, nw = (blockDim.x + 31) >> 5;
  const bool live = h < H;
  const float* src = theta + (int64_t)row * ld;
  double w1[DP], w2[OP], b2[OP], b1 = 0.0;
#pragma unroll
  for (int d = 0; d < DP; ++d) w1[d] = (live && d < D) ? (double)src[d * H + h] : 0.0;
#pragma unroll
  for (int o = 0; o < OP; ++o) {
    w2[o] = (live && o < O) ? (double)src[D * H + h * O + o] : 0.0;
    b2[o] = (o < O) ? (double)src[D * H + H * O + H + o] : 0.0;
  }
  if (live) b1 = (double)src[D * H + H * O + h];
  for (int d = h; d < DP; d += blockDim.x) xs[0][d] = (d < D && n > 0) ? (double)x[d] : 0.0;
  if (h == 0 && n > 0) labs[0] = y[0];
  __syncthreads();
  for (int64_t r = 0; r < n; ++r) {
    const int cur = (int)(r % 3), nxt = (int)((r + 1) % 3), par = (int)(r & 1);
    // issue the next row's loads now, park them in shared memory just before the barrier: the global-load
    // latency overlaps the forward pass instead of stalling it
    float xn[(DP + 31) / 32];
    int labn = 0;
    if (r + 1 < n) {
#pragma unroll
      for (int q = 0; q < (DP + 31) / 32; ++q) { const int d = h + 32 * q; xn[q] = (h < 32 && d < D) ? __ldg(x + (r + 1) * xstride + d) : 0.f; }
      if (h == 0) labn = __ldg(y + r + 1);
    }
    double z4[4] = {0.0, 0.0, 0.0, 0.0};                                   // four partial sums: the FMA chain is latency bound
#pragma unroll
    for (int d = 0; d < DP; ++d) z4[d & 3] += xs[cur][d] * w1[d];          // LG:70 (padding columns are zero)
    const double z = ((z4[0] + z4[1]) + (z4[2] + z4[3])) - b1;
    const double hid = live ? 1.0 / (1.0 + exp(-z)) : 0.0;                 // LG:71
#pragma unroll
    for (int o = 0; o < OP; ++o) {
      double pz = hid * w2[o];
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) pz += __shfl_xor_sync(0xffffffffu, pz, s);
      if (lane == 0) red[par][warp][o] = pz;
    }
    if (r + 1 < n && h < 32) {
#pragma unroll
      for (int q = 0; q < (DP + 31) / 32; ++q) { const int d = h + 32 * q; if (d < DP) xs[nxt][d] = (double)xn[q]; }
      if (h == 0) labs[nxt] = labn;
    }
    __syncthreads();
    const int label = labs[cur];
    // each warp evaluates the O outputs once, lane o doing output o (the fp64 exp + divide chain is the latency of
    // this loop; doing the O of them side by side instead of one after the other is what matters), then broadcasts
    double my_od = 0.0;
    double od[OP], back = 0.0;
    {
      // lane-parallel: lane o owns b2[o] through a select chain (b2 is replicated in registers)
      double zsum = 0.0, bsel = 0.0;
#pragma unroll
      for (int o = 0; o < OP; ++o) {
        double t = 0.0;
        for (int w = 0; w < nw; ++w) t += red[par][w][o];
        if (lane == o) { zsum = t; bsel = b2[o]; }
      }
      const double z2 = zsum - bsel;
      const double out = 1.0 / (1.0 + exp(-z2));                           // LG:73
      const double desired = (lane == label) ? 1.0 : 0.0;
      my_od = (lane < O) ? (desired - out) * (out * (1.0 - out)) : 0.0;    // LG:100
#pragma unroll
      for (int o = 0; o < OP; ++o) {
        od[o] = __shfl_sync(0xffffffffu, my_od, o);
        back += od[o] * w2[o];                                             // LG:101 (W2 BEFORE its update)
      }
    }
    const double hd = back * (hid * (1.0 - hid));
#pragma unroll
    for (int o = 0; o < OP; ++o) {
      w2[o] += hid * (od[o] * lr);                                         // LG:102
      b2[o] += (-1.0 * lr) * od[o];                                        // LG:103
    }
    const double hl = hd * lr;
#pragma unroll
    for (int d = 0; d < DP; ++d) w1[d] += xs[cur][d] * hl;                 // LG:105
    b1 += (-1.0 * lr) * hd;                                                // LG:106
  }
  float* dst = theta_gd + (int64_t)row * ld;
  if (live) {
#pragma unroll
    for (int d = 0; d < DP; ++d) if (d < D) dst[d * H + h] = (float)w1[d];
#pragma unroll
    for (int o = 0; o < OP; ++o) if (o < O) dst[D * H + h * O + o] = (float)w2[o];
    dst[D * H + H * O + h] = (float)b1;
  }
  if (h == 0) {
#pragma unroll
    for (int o = 0; o < OP; ++o) if (o < O) dst[D * H + H * O + H + o] = (float)b2[o];
  }
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

// ---- batch mode, register-tiled version (HP <= 64): the production path for large N ---------------------------
// Phase A (thread = row, features in registers, weights via broadcast LDS.128): forward pass, output deltas,
// hidden deltas; the tile (x, hid, hd, od) is parked in shared memory. Phase B: the outer products
// X^T.hd (D x H), hid^T.od (H x O) and the bias sums are accumulated over the 128 rows by threads that own fixed
// 4x4 blocks of the gradient in REGISTERS for the whole row chunk (2 FMA-instructions per shared-memory wavefront).
template <int DP, int OP, int NB>
__global__ void __launch_bounds__(kBT) langevin_batch_tiled_kernel(const float* x, const int32_t* y, int64_t n, int D, int H, int O, int HP,
                                                                   const float* theta, int64_t ld, const int32_t* rows,
                                                                   const int32_t* n_rows, int R, int tiles_per_chunk, float* part) {
  const int nact = n_rows ? *n_rows : R;
  if ((int)blockIdx.y >= nact) return;
  const int row = rows ? rows[blockIdx.y] : (int)blockIdx.y;
  extern __shared__ __align__(16) float smf[];
  const int P = D * H + H * O + H + O;
  const int HP4 = HP >> 2;
  float* sW1 = smf;                    // [DP][HP]
  float* sW2 = sW1 + DP * HP;          // [HP][OP]
  float* sB1 = sW2 + HP * OP;          // [HP]
  float* sB2 = sB1 + HP;               // [OP]
  float* xt = sB2 + OP;                // [kBT][DP]
  float* ht = xt + kBT * DP;           // [kBT][HP]
  float* hd = ht + kBT * HP;           // [kBT][HP]
  float* od = hd + kBT * HP;           // [kBT][OP]
  const int tid = threadIdx.x;
  const float* src = theta + (int64_t)row * ld;
  for (int i = tid; i < DP * HP; i += kBT) { const int d = i / HP, h = i - d * HP; sW1[i] = (d < D && h < H) ? src[d * H + h] : 0.f; }
  for (int i = tid; i < HP * OP; i += kBT) { const int h = i / OP, o = i - h * OP; sW2[i] = (h < H && o < O) ? src[D * H + h * O + o] : 0.f; }
  for (int i = tid; i < HP; i += kBT) sB1[i] = (i < H) ? src[D * H + H * O + i] : 0.f;
  for (int i = tid; i < OP; i += kBT) sB2[i] = (i < O) ? src[D * H + H * O + H + i] : 0.f;
  float acc1[NB][16];
  float acc2[16];
#pragma unroll
  for (int b = 0; b < NB; ++b)
#pragma unroll
    for (int i = 0; i < 16; ++i) acc1[b][i] = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) acc2[i] = 0.f;
  float accb = 0.f;
  const int nblk1 = (DP / 4) * HP4, nblk2 = HP4 * (OP / 4);
  __syncthreads();
  const int64_t n_tiles = (n + kBT - 1) / kBT;
  const int64_t t0 = (int64_t)blockIdx.x * tiles_per_chunk, t1 = min(t0 + tiles_per_chunk, n_tiles);
  for (int64_t t = t0; t < t1; ++t) {
    // ---------------- phase A ----------------
    const int64_t r = t * kBT + tid;
    const bool valid = r < n;
    float xr[DP];
    {
      const float4* xp = reinterpret_cast<const float4*>(x + (valid ? r : 0) * DP);
#pragma unroll
      for (int d4 = 0; d4 < DP / 4; ++d4) {
        float4 v = valid ? __ldg(xp + d4) : make_float4(0.f, 0.f, 0.f, 0.f);
        xr[4 * d4] = v.x; xr[4 * d4 + 1] = v.y; xr[4 * d4 + 2] = v.z; xr[4 * d4 + 3] = v.w;
        *reinterpret_cast<float4*>(xt + tid * DP + 4 * d4) = v;
      }
    }
    const int label = valid ? y[r] : -1;
    float out[OP];
#pragma unroll
    for (int o = 0; o < OP; ++o) out[o] = -sB2[o];
    for (int h4 = 0; h4 < HP; h4 += 4) {
      const float4 b = *reinterpret_cast<const float4*>(sB1 + h4);
      float z0 = -b.x, z1 = -b.y, z2 = -b.z, z3 = -b.w;
#pragma unroll
      for (int d = 0; d < DP; ++d) {
        const float4 w = *reinterpret_cast<const float4*>(sW1 + d * HP + h4);
        z0 = fmaf(xr[d], w.x, z0); z1 = fmaf(xr[d], w.y, z1); z2 = fmaf(xr[d], w.z, z2); z3 = fmaf(xr[d], w.w, z3);
      }
      float hv[4] = {1.f / (1.f + __expf(-z0)), 1.f / (1.f + __expf(-z1)), 1.f / (1.f + __expf(-z2)), 1.f / (1.f + __expf(-z3))};
      *reinterpret_cast<float4*>(ht + tid * HP + h4) = make_float4(hv[0], hv[1], hv[2], hv[3]);
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int o4 = 0; o4 < OP; o4 += 4) {
          const float4 w = *reinterpret_cast<const float4*>(sW2 + (h4 + j) * OP + o4);
          out[o4] = fmaf(hv[j], w.x, out[o4]); out[o4 + 1] = fmaf(hv[j], w.y, out[o4 + 1]);
          out[o4 + 2] = fmaf(hv[j], w.z, out[o4 + 2]); out[o4 + 3] = fmaf(hv[j], w.w, out[o4 + 3]);
        }
    }
    float odl[OP];
#pragma unroll
    for (int o = 0; o < OP; ++o) {
      const float s = 1.f / (1.f + __expf(-out[o]));
      const float desired = (o == label) ? 1.f : 0.f;
      odl[o] = (valid && o < O) ? (desired - s) * (s * (1.f - s)) : 0.f;   // LG:100
    }
#pragma unroll
    for (int o4 = 0; o4 < OP; o4 += 4) *reinterpret_cast<float4*>(od + tid * OP + o4) = make_float4(odl[o4], odl[o4 + 1], odl[o4 + 2], odl[o4 + 3]);
    for (int h4 = 0; h4 < HP; h4 += 4) {
      const float4 hvv = *reinterpret_cast<const float4*>(ht + tid * HP + h4);
      const float hv[4] = {hvv.x, hvv.y, hvv.z, hvv.w};
      float dl[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float sacc = 0.f;
#pragma unroll
        for (int o4 = 0; o4 < OP; o4 += 4) {
          const float4 w = *reinterpret_cast<const float4*>(sW2 + (h4 + j) * OP + o4);
          sacc = fmaf(odl[o4], w.x, sacc); sacc = fmaf(odl[o4 + 1], w.y, sacc); sacc = fmaf(odl[o4 + 2], w.z, sacc); sacc = fmaf(odl[o4 + 3], w.w, sacc);
        }
        dl[j] = sacc * (hv[j] * (1.f - hv[j]));                               // LG:101
      }
      *reinterpret_cast<float4*>(hd + tid * HP + h4) = make_float4(dl[0], dl[1], dl[2], dl[3]);
    }
    __syncthreads();
    // ---------------- phase B ----------------
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const int blk = tid + b * kBT;
      if (blk < nblk1) {
        const int db = blk / HP4, hb = blk - db * HP4;
        const float* xp = xt + db * 4;
        const float* hp = hd + hb * 4;
#pragma unroll 4
        for (int rr = 0; rr < kBT; ++rr) {
          const float4 xv = *reinterpret_cast<const float4*>(xp + rr * DP);
          const float4 hv = *reinterpret_cast<const float4*>(hp + rr * HP);
          const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
          const float ha[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc1[b][4 * i + j] = fmaf(xa[i], ha[j], acc1[b][4 * i + j]);
        }
      }
    }
    if (tid < nblk2) {
      const int hb = tid / (OP / 4), ob = tid - hb * (OP / 4);
      const float* hp = ht + hb * 4;
      const float* op = od + ob * 4;
#pragma unroll 4
      for (int rr = 0; rr < kBT; ++rr) {
        const float4 hv = *reinterpret_cast<const float4*>(hp + rr * HP);
        const float4 ov = *reinterpret_cast<const float4*>(op + rr * OP);
        const float ha[4] = {hv.x, hv.y, hv.z, hv.w};
        const float oa[4] = {ov.x, ov.y, ov.z, ov.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc2[4 * i + j] = fmaf(ha[i], oa[j], acc2[4 * i + j]);
      }
    } else if (tid >= kBT - HP - OP) {   // the last HP + OP threads own the bias sums (requires nblk2 + HP + OP <= kBT)
      const int q = tid - (kBT - HP - OP);
      const float* colp = (q < HP) ? (hd + q) : (od + (q - HP));
      const int stride = (q < HP) ? HP : OP;
      float sacc = 0.f;
      for (int rr = 0; rr < kBT; ++rr) sacc += colp[rr * stride];
      accb -= sacc;                                                            // LG:103,106: biases move by -lr*delta
    }
    __syncthreads();
  }
  float* dst = part + ((int64_t)blockIdx.x * R + blockIdx.y) * P;
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const int blk = tid + b * kBT;
    if (blk < nblk1) {
      const int db = blk / HP4, hb = blk - db * HP4;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int d = db * 4 + i, h = hb * 4 + j;
          if (d < D && h < H) dst[d * H + h] = acc1[b][4 * i + j];
        }
    }
  }
  if (tid < nblk2) {
    const int hb = tid / (OP / 4), ob = tid - hb * (OP / 4);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int h = hb * 4 + i, o = ob * 4 + j;
        if (h < H && o < O) dst[D * H + h * O + o] = acc2[4 * i + j];
      }
  } else if (tid >= kBT - HP - OP) {
    const int q = tid - (kBT - HP - OP);
    if (q < HP) { if (q < H) dst[D * H + H * O + q] = accb; }
    else if (q - HP < O) dst[D * H + H * O + H + (q - HP)] = accb;
  }
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

int launch_langevin_batch_finalize(const float* part, int n_chunks, int R, int P, const float* theta, float* theta_gd, int64_t ld,
                                   const int32_t* rows, const int32_t* n_rows, double lr_scale, cudaStream_t st) {
  langevin_batch_finalize<<<dim3((P + 127) / 128, R), 128, 0, st>>>(part, n_chunks, R, P, theta, theta_gd, ld, rows, n_rows, lr_scale);
  SAPT_CHECK_LAUNCH();
  return SAPT_OK;
}

bool langevin_tc_supported(const sapt_problem* p);   // sapt_langevin_tc.cu
int launch_langevin_tc(sapt_problem* p, const float* theta, float* theta_gd, int64_t ld, int R, const int32_t* rows, const int32_t* n_rows,
                       double lr_scale, cudaStream_t st);

}  // namespace sapt

using namespace sapt;

extern "C" int sapt_langevin(sapt_problem_t* p, const float* theta, float* theta_gd, int64_t ld, int R, const int32_t* rows,
                             const int32_t* n_rows, double lrate, int mode, double batch_scale, void* stream) {
  SAPT_REQUIRE(p && theta && theta_gd && R >= 0 && ld >= p->P, "sapt_langevin");
  if (R == 0) return SAPT_OK;
  const DataSet& ds = p->set[0];
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 0) {
    if (p->H <= 128) {
      const int threads = ((p->H + 31) / 32) * 32;
#define SAPT_SGD(DPV, OPV) langevin_sgd_reg_kernel<DPV, OPV><<<R, threads, 0, st>>>(ds.x, p->DP, ds.y, ds.n, p->D, p->H, p->O, theta, theta_gd, ld, rows, n_rows, R, lrate);
#define SAPT_SGD_OP(DPV)                                                          \
  switch (p->OP) {                                                                \
    case 4: SAPT_SGD(DPV, 4) break;                                               \
    case 8: SAPT_SGD(DPV, 8) break;                                               \
    case 12: SAPT_SGD(DPV, 12) break;                                             \
    case 16: SAPT_SGD(DPV, 16) break;                                             \
    case 20: SAPT_SGD(DPV, 20) break;                                             \
    default: set_error("sapt_langevin: O padded to %d unsupported", p->OP); return SAPT_ERR_UNSUPPORTED; \
  }
      switch ((p->D + 7) / 8) {  // features padded to a multiple of 8 only: the register-resident column is D8 doubles
        case 1: SAPT_SGD_OP(8) break;
        case 2: SAPT_SGD_OP(16) break;
        case 3: SAPT_SGD_OP(24) break;
        case 4: SAPT_SGD_OP(32) break;
        case 5: SAPT_SGD_OP(40) break;
        case 6: SAPT_SGD_OP(48) break;
        case 7: SAPT_SGD_OP(56) break;
        case 8: SAPT_SGD_OP(64) break;
        default: set_error("sapt_langevin: D = %d unsupported", p->D); return SAPT_ERR_UNSUPPORTED;
      }
#undef SAPT_SGD_OP
#undef SAPT_SGD
      SAPT_CHECK_LAUNCH();
      return SAPT_OK;
    }
    const size_t smem = (size_t)(p->P + p->D + 2 * p->H + p->O) * sizeof(double);
    if (smem > 200 * 1024) { set_error("sapt_langevin: topology too large for the sgd_exact kernel"); return SAPT_ERR_UNSUPPORTED; }
    if (smem > 48 * 1024) SAPT_CUDA(cudaFuncSetAttribute(langevin_sgd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    langevin_sgd_kernel<<<R, 256, smem, st>>>(ds.x, ds.y, ds.n, p->DP, p->D, p->H, p->O, theta, theta_gd, ld, rows, n_rows, R, lrate);
    SAPT_CHECK_LAUNCH();
    return SAPT_OK;
  }
  if (mode == 1 && langevin_tc_supported(p)) return launch_langevin_tc(p, theta, theta_gd, ld, R, rows, n_rows, lrate * batch_scale, st);
  if (mode == 1 || mode == 2) {   // 2: the CUDA-core kernels even where the tensor-core path applies
    const int P = p->P;
    const int64_t n_tiles = (ds.n + kBT - 1) / kBT;
    int64_t n_chunks = std::min<int64_t>(n_tiles, std::max<int64_t>(1, ((int64_t)p->num_sms * 4 + R - 1) / R));
    const int tiles_per_chunk = (int)((n_tiles + n_chunks - 1) / n_chunks);
    n_chunks = (n_tiles + tiles_per_chunk - 1) / tiles_per_chunk;
    int rc = ensure_workspace(p, (size_t)n_chunks * R * P * sizeof(float));
    if (rc) return rc;
    const dim3 grid((unsigned)n_chunks, (unsigned)R);
    const int HP = p->HP, OP = p->OP, DP = p->DP;
    const bool tiled = HP <= 64 && (HP / 4) * (OP / 4) + HP + OP <= kBT;
    if (tiled) {
      const size_t smem = (size_t)(DP * HP + HP * OP + HP + OP + kBT * (DP + 2 * HP + OP)) * sizeof(float);
#define SAPT_LB(DPV, OPV, NBV)                                                                                            \
  {                                                                                                                       \
    auto k = langevin_batch_tiled_kernel<DPV, OPV, NBV>;                                                                  \
    if (smem > 48 * 1024) SAPT_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    k<<<grid, kBT, smem, st>>>(ds.x, ds.y, ds.n, p->D, p->H, p->O, HP, theta, ld, rows, n_rows, R, tiles_per_chunk, (float*)p->ws); \
  }
#define SAPT_LB_OP(DPV, NBV)                                                          \
  switch (OP) {                                                                       \
    case 4: SAPT_LB(DPV, 4, NBV) break;                                               \
    case 8: SAPT_LB(DPV, 8, NBV) break;                                               \
    case 12: SAPT_LB(DPV, 12, NBV) break;                                             \
    case 16: SAPT_LB(DPV, 16, NBV) break;                                             \
    case 20: SAPT_LB(DPV, 20, NBV) break;                                             \
    default: set_error("sapt_langevin: O padded to %d unsupported", OP); return SAPT_ERR_UNSUPPORTED; \
  }
      switch (DP) {
        case 8: SAPT_LB_OP(8, 1) break;
        case 16: SAPT_LB_OP(16, 1) break;
        case 32: SAPT_LB_OP(32, 1) break;
        case 64: SAPT_LB_OP(64, 2) break;
        default: set_error("sapt_langevin: D padded to %d unsupported", DP); return SAPT_ERR_UNSUPPORTED;
      }
#undef SAPT_LB_OP
#undef SAPT_LB
      SAPT_CHECK_LAUNCH();
    } else {
      const size_t smem = (size_t)(2 * P + kBT * ((p->D | 1) + 2 * (p->H | 1) + (p->O | 1))) * sizeof(float);
      if (smem > 220 * 1024) { set_error("sapt_langevin: topology too large for the batch kernel"); return SAPT_ERR_UNSUPPORTED; }
      if (smem > 48 * 1024) SAPT_CUDA(cudaFuncSetAttribute(langevin_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      langevin_batch_kernel<<<grid, kBT, smem, st>>>(ds.x, ds.y, ds.n, p->DP, p->D, p->H, p->O, theta, ld, rows, n_rows, R, tiles_per_chunk,
                                                     (float*)p->ws);
      SAPT_CHECK_LAUNCH();
    }
    langevin_batch_finalize<<<dim3((P + 127) / 128, R), 128, 0, st>>>((const float*)p->ws, (int)n_chunks, R, P, theta, theta_gd, ld, rows, n_rows,
                                                                      lrate * batch_scale);
    SAPT_CHECK_LAUNCH();
    return SAPT_OK;
  }
  set_error("sapt_langevin: unknown mode %d", mode);
  return SAPT_ERR_INVALID;
}
