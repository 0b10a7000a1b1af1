// fp32 evaluation of one chain on the rows a thread owns — shared by the tiled likelihood kernel
// (sapt_loglik_f32.cu) and the fused small-problem step kernel (sapt_fused.cu).
#pragma once
#include "sapt_device.cuh"

namespace sapt {

// stage one chain's weights in shared memory, zero padded to [DP][HP] / [HP][OP] and pre-scaled by -log2(e)
// (biases by +log2(e): z = x.W - B), so that a sigmoid is ex2 + add + rcp
template <int DP, int OP>
__device__ __forceinline__ void stage_chain_weights(const float* __restrict__ src, float* dst, int D, int H, int O, int HP, int tid, int nthreads) {
  const int n1 = DP * HP, n2 = HP * OP;
  const int S = n1 + n2 + HP + OP;
  for (int i = tid; i < S; i += nthreads) {
    float v = 0.f;
    if (i < n1) {
      const int d = i / HP, h = i - d * HP;
      if (d < D && h < H) v = -kLog2e * src[d * H + h];
    } else if (i < n1 + n2) {
      const int j = i - n1;
      const int h = j / OP, o = j - h * OP;
      if (h < H && o < O) v = -kLog2e * src[D * H + h * O + o];
    } else if (i < n1 + n2 + HP) {
      const int h = i - n1 - n2;
      if (h < H) v = kLog2e * src[D * H + H * O + h];
    } else {
      const int o = i - n1 - n2 - HP;
      if (o < O) v = kLog2e * src[D * H + H * O + H + o];
    }
    dst[i] = v;
  }
}

// one chain, RPT rows of this thread
template <int DP, int OP, int RPT>
__device__ __forceinline__ void eval_rows(const float* __restrict__ sw, int HP, int O, const float (&x)[RPT][DP],
                                          const int (&y)[RPT], float (&lp)[RPT], int (&pred)[RPT]) {
  const float* sW1 = sw;
  const float* sW2 = sw + DP * HP;
  const float* sB1 = sW2 + HP * OP;
  const float* sB2 = sB1 + HP;
  float out[RPT][OP];
#pragma unroll
  for (int o4 = 0; o4 < OP; o4 += 4) {
    const float4 b = *reinterpret_cast<const float4*>(sB2 + o4);
#pragma unroll
    for (int r = 0; r < RPT; ++r) { out[r][o4] = b.x; out[r][o4 + 1] = b.y; out[r][o4 + 2] = b.z; out[r][o4 + 3] = b.w; }
  }
  for (int h4 = 0; h4 < HP; h4 += 4) {
    float z[RPT][4];
    {
      const float4 b = *reinterpret_cast<const float4*>(sB1 + h4);
#pragma unroll
      for (int r = 0; r < RPT; ++r) { z[r][0] = b.x; z[r][1] = b.y; z[r][2] = b.z; z[r][3] = b.w; }
    }
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      const float4 w = *reinterpret_cast<const float4*>(sW1 + d * HP + h4);
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        z[r][0] = fmaf(x[r][d], w.x, z[r][0]);
        z[r][1] = fmaf(x[r][d], w.y, z[r][1]);
        z[r][2] = fmaf(x[r][d], w.z, z[r][2]);
        z[r][3] = fmaf(x[r][d], w.w, z[r][3]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float hv[RPT];
#pragma unroll
      for (int r = 0; r < RPT; ++r) hv[r] = sigmoid_from_neg_log2e(z[r][j]);
#pragma unroll
      for (int o4 = 0; o4 < OP; o4 += 4) {
        const float4 w = *reinterpret_cast<const float4*>(sW2 + (h4 + j) * OP + o4);
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
          out[r][o4] = fmaf(hv[r], w.x, out[r][o4]);
          out[r][o4 + 1] = fmaf(hv[r], w.y, out[r][o4 + 1]);
          out[r][o4 + 2] = fmaf(hv[r], w.z, out[r][o4 + 2]);
          out[r][o4 + 3] = fmaf(hv[r], w.w, out[r][o4 + 3]);
        }
      }
    }
  }
  // out[] holds t = -log2e * z2. argmax(sigmoid(z2)) == argmin(t), first minimum on ties (np.argmax).
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    float tmin = out[r][0];
    int pm = 0;
    float sum = 0.f, oy = 0.f;
#pragma unroll
    for (int o = 0; o < OP; ++o) {
      if (o < O) {
        const float t = out[r][o];
        if (t < tmin) { tmin = t; pm = o; }
        const float s = sigmoid_from_neg_log2e(t);
        sum += ex2_approx(s * kLog2e);
        if (o == y[r]) oy = s;
      }
    }
    lp[r] = oy - kLn2 * lg2_approx(sum);  // log softmax(out)[y]
    pred[r] = pm;
  }
}


}  // namespace sapt
