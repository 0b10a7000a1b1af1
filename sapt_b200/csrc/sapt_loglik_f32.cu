// fp32 CUDA-core tiles of the batched likelihood (the precision-safe path and the fallback
// for topologies outside the tcgen05 envelope).
//
// Replaces, for R chains at once: ptReplica.likelihood_func (_revamp.py:391-402) ->
// Network.evaluate_proposal (:155-174) -> ForwardPass (:70-76) + softmax (:129-131).
//
// Decomposition: grid = (row chunks, replica groups). A CTA stages the weights of up to G
// chains in shared memory (pre-scaled by -log2 e so a sigmoid is ex2+add+rcp), then walks its
// row chunk tile by tile; a thread owns RPT rows whose features stay in registers while the
// G chains are evaluated against them, weights arriving as broadcast LDS.128. Per-thread
// log-likelihood sums are kept in fp64 (one DADD per ~1000 instructions); the chunk partials
// are reduced across the CTA with warp shuffles and written once per (chunk, chain); a small
// second kernel folds the chunks in fp64.
#include <algorithm>

#include "sapt_device.cuh"
#include "sapt_eval_f32.cuh"
#include "sapt_internal.h"

namespace sapt {

constexpr int kThreads = 128;
constexpr int kGMax = 8;

struct LoglikF32Params {
  const float* x; const int32_t* y; int64_t n_pad;
  const float* theta; int64_t ld;
  const int32_t* rows; const int32_t* n_rows; int R;
  int D, H, O, HP;
  int G; int tiles_per_chunk; int n_chunks;
  double* partL; int32_t* partNc; int32_t* partSq;  // [n_chunks][R] (indexed by position in the active list)
};

template <int DP, int OP, int RPT>
__global__ void __launch_bounds__(kThreads) loglik_f32_kernel(LoglikF32Params p) {
  extern __shared__ __align__(16) float smem[];
  __shared__ double red_d[kThreads / 32];
  __shared__ int red_i[2][kThreads / 32];
  const int tid = threadIdx.x;
  const int nact = p.n_rows ? *p.n_rows : p.R;
  const int g0 = blockIdx.y * p.G;
  if (g0 >= nact) return;
  const int ng = min(p.G, nact - g0);
  const int HP = p.HP, D = p.D, H = p.H, O = p.O;
  const int S = DP * HP + HP * OP + HP + OP;  // floats per staged chain

  // ---- stage weights, scaled by -log2(e) (biases by +log2(e): z = x.W - B) ----
  for (int g = 0; g < ng; ++g) {
    const int row = p.rows ? p.rows[g0 + g] : (g0 + g);
    stage_chain_weights<DP, OP>(p.theta + (int64_t)row * p.ld, smem + g * S, D, H, O, HP, tid, kThreads);
  }
  __syncthreads();

  // per-thread, per-chain accumulators live in shared memory (fp64 log-likelihood, exact ints):
  // one LDS/STS pair per ~1000 instructions of evaluation, and the chain loop stays rolled.
  double* accL = reinterpret_cast<double*>(smem + (size_t)p.G * S);  // [G][kThreads]
  int* accNc = reinterpret_cast<int*>(accL + (size_t)p.G * kThreads);
  int* accSq = accNc + (size_t)p.G * kThreads;
  for (int g = 0; g < ng; ++g) { accL[g * kThreads + tid] = 0.0; accNc[g * kThreads + tid] = 0; accSq[g * kThreads + tid] = 0; }

  const int64_t tile_rows = (int64_t)kThreads * RPT;
  const int64_t t_begin = (int64_t)blockIdx.x * p.tiles_per_chunk;
  const int64_t t_end = min(t_begin + p.tiles_per_chunk, (p.n_pad + tile_rows - 1) / tile_rows);
  for (int64_t t = t_begin; t < t_end; ++t) {
    float x[RPT][DP];
    int y[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int64_t row = t * tile_rows + (int64_t)r * kThreads + tid;
      if (row < p.n_pad) {
        const float4* xr = reinterpret_cast<const float4*>(p.x + row * DP);
#pragma unroll
        for (int d4 = 0; d4 < DP / 4; ++d4) {
          const float4 v = __ldg(xr + d4);
          x[r][4 * d4] = v.x; x[r][4 * d4 + 1] = v.y; x[r][4 * d4 + 2] = v.z; x[r][4 * d4 + 3] = v.w;
        }
        y[r] = __ldg(p.y + row);
      } else {
#pragma unroll
        for (int d = 0; d < DP; ++d) x[r][d] = 0.f;
        y[r] = -1;
      }
    }
#pragma unroll 1
    for (int g = 0; g < ng; ++g) {
      float lp[RPT];
      int pred[RPT];
      eval_rows<DP, OP, RPT>(smem + g * S, HP, O, x, y, lp, pred);
      double l = 0.0;
      int a = 0, b = 0;
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        if (y[r] >= 0) {
          l += (double)lp[r];
          a += (pred[r] == y[r]);
          const int e = pred[r] - y[r];
          b += e * e;
        }
      }
      accL[g * kThreads + tid] += l;
      accNc[g * kThreads + tid] += a;
      accSq[g * kThreads + tid] += b;
    }
  }

  // ---- chunk partials: warp shuffle, then one smem hop ----
  const int lane = tid & 31, warp = tid >> 5;
  for (int g = 0; g < ng; ++g) {
    const double l = warp_sum(accL[g * kThreads + tid]);
    const int a = warp_sum(accNc[g * kThreads + tid]);
    const int b = warp_sum(accSq[g * kThreads + tid]);
    __syncthreads();
    if (lane == 0) { red_d[warp] = l; red_i[0][warp] = a; red_i[1][warp] = b; }
    __syncthreads();
    if (tid == 0) {
      double ls = 0; int as = 0, bs = 0;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) { ls += red_d[w]; as += red_i[0][w]; bs += red_i[1][w]; }
      const int64_t o = (int64_t)blockIdx.x * p.R + (g0 + g);
      p.partL[o] = ls; p.partNc[o] = as; p.partSq[o] = bs;
    }
  }
}

// folds the chunk partials (fp64) and scatters to the theta-row-indexed outputs
__global__ void loglik_finalize_kernel(const double* partL, const int32_t* partNc, const int32_t* partSq, int n_chunks, int R,
                                       const int32_t* rows, const int32_t* n_rows, double* L, int32_t* nc, double* sq) {
  const int nact = n_rows ? *n_rows : R;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= nact) return;
  double l = 0; long long a = 0, b = 0;
  for (int c = lane; c < n_chunks; c += 32) {
    const int64_t o = (int64_t)c * R + warp;
    l += partL[o]; a += partNc[o]; b += partSq[o];
  }
  l = warp_sum(l); a = warp_sum(a); b = warp_sum(b);
  if (lane == 0) {
    const int row = rows ? rows[warp] : warp;
    L[row] = l;
    if (nc) nc[row] = (int32_t)a;
    if (sq) sq[row] = (double)b;
  }
}

// per-row outputs for ONE theta: fx = argmax class, prob = softmax over the sigmoid outputs
template <int DP, int OP>
__global__ void __launch_bounds__(kThreads) forward_rows_kernel(LoglikF32Params p, int64_t n, int32_t* fx, float* prob) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int HP = p.HP, D = p.D, H = p.H, O = p.O;
  const int S = DP * HP + HP * OP + HP + OP;
  const int n1 = DP * HP, n2 = HP * OP;
  for (int i = tid; i < S; i += kThreads) {
    float v = 0.f;
    if (i < n1) { const int d = i / HP, h = i - d * HP; if (d < D && h < H) v = -kLog2e * p.theta[d * H + h]; }
    else if (i < n1 + n2) { const int j = i - n1; const int h = j / OP, o = j - h * OP; if (h < H && o < O) v = -kLog2e * p.theta[D * H + h * O + o]; }
    else if (i < n1 + n2 + HP) { const int h = i - n1 - n2; if (h < H) v = kLog2e * p.theta[D * H + H * O + h]; }
    else { const int o = i - n1 - n2 - HP; if (o < O) v = kLog2e * p.theta[D * H + H * O + H + o]; }
    smem[i] = v;
  }
  __syncthreads();
  const int64_t row = (int64_t)blockIdx.x * kThreads + tid;
  if (row >= n) return;
  const float* sW1 = smem; const float* sW2 = smem + n1; const float* sB1 = sW2 + n2; const float* sB2 = sB1 + HP;
  float x[DP];
  const float4* xr = reinterpret_cast<const float4*>(p.x + row * DP);
#pragma unroll
  for (int d4 = 0; d4 < DP / 4; ++d4) { const float4 v = __ldg(xr + d4); x[4 * d4] = v.x; x[4 * d4 + 1] = v.y; x[4 * d4 + 2] = v.z; x[4 * d4 + 3] = v.w; }
  float out[OP];
#pragma unroll
  for (int o = 0; o < OP; ++o) out[o] = sB2[o];
  for (int h = 0; h < HP; ++h) {
    float z = sB1[h];
#pragma unroll
    for (int d = 0; d < DP; ++d) z = fmaf(x[d], sW1[d * HP + h], z);
    const float hv = sigmoid_from_neg_log2e(z);
#pragma unroll
    for (int o = 0; o < OP; ++o) out[o] = fmaf(hv, sW2[h * OP + o], out[o]);
  }
  float tmin = out[0]; int pm = 0; float sum = 0.f; float e[OP];
#pragma unroll
  for (int o = 0; o < OP; ++o) {
    e[o] = 0.f;
    if (o < O) {
      if (out[o] < tmin) { tmin = out[o]; pm = o; }
      e[o] = expf(1.0f / (1.0f + exp2f(out[o])));
      sum += e[o];
    }
  }
  fx[row] = pm;
#pragma unroll
  for (int o = 0; o < OP; ++o) if (o < O) prob[row * O + o] = e[o] / sum;
}

template <int DP, int OP, int RPT>
static int launch_one(const LoglikF32Params& p, dim3 grid, size_t smem, cudaStream_t st) {
  auto k = loglik_f32_kernel<DP, OP, RPT>;
  if (smem > 48 * 1024) SAPT_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<grid, kThreads, smem, st>>>(p);
  SAPT_CHECK_LAUNCH();
  return SAPT_OK;
}

template <int DP, int RPT>
static int dispatch_op(int OP, const LoglikF32Params& p, dim3 grid, size_t smem, cudaStream_t st) {
  switch (OP) {
    case 4: return launch_one<DP, 4, RPT>(p, grid, smem, st);
    case 8: return launch_one<DP, 8, RPT>(p, grid, smem, st);
    case 12: return launch_one<DP, 12, RPT>(p, grid, smem, st);
    case 16: return launch_one<DP, 16, RPT>(p, grid, smem, st);
    case 20: return launch_one<DP, 20, RPT>(p, grid, smem, st);
    default: set_error("fp32 path: O padded to %d unsupported (O <= 20)", OP); return SAPT_ERR_UNSUPPORTED;
  }
}

static int rpt_for(int DP) { return DP <= 16 ? 2 : 1; }

int launch_loglik_f32(sapt_problem* p, int which, const float* theta, int64_t ld, int R, const int32_t* rows,
                      const int32_t* n_rows, double* L, int32_t* nc, double* sq, cudaStream_t st) {
  const DataSet& ds = p->set[which];
  if (R <= 0 || ds.n <= 0) return SAPT_OK;
  LoglikF32Params a{};
  a.x = ds.x; a.y = ds.y; a.n_pad = ds.n_pad; a.theta = theta; a.ld = ld; a.rows = rows; a.n_rows = n_rows; a.R = R;
  a.D = p->D; a.H = p->H; a.O = p->O; a.HP = p->HP;
  const int S = p->DP * p->HP + p->HP * p->OP + p->HP + p->OP;
  const size_t per_chain_w = (size_t)S * sizeof(float);
  const size_t per_chain = ((per_chain_w + 7) / 8) * 8 + (size_t)kThreads * (sizeof(double) + 2 * sizeof(int));
  int G = (int)std::min<size_t>(kGMax, (56 * 1024) / per_chain);
  if (G < 1) G = 1;
  if (per_chain > 200 * 1024) { set_error("fp32 path: topology needs %zu B of shared memory per chain", per_chain); return SAPT_ERR_UNSUPPORTED; }
  const int rpt = rpt_for(p->DP);
  const int64_t tile_rows = (int64_t)kThreads * rpt;
  const int64_t n_tiles = (ds.n_pad + tile_rows - 1) / tile_rows;
  // few chains and few rows: spread chains over more CTAs
  while (G > 1 && (int64_t)((R + G - 1) / G) * n_tiles < 2 * p->num_sms) G >>= 1;
  const int n_groups = (R + G - 1) / G;
  int64_t want_chunks = std::max<int64_t>(1, ((int64_t)p->num_sms * 8 + n_groups - 1) / n_groups);
  int64_t n_chunks = std::min<int64_t>(want_chunks, n_tiles);
  const int tiles_per_chunk = (int)((n_tiles + n_chunks - 1) / n_chunks);
  n_chunks = (n_tiles + tiles_per_chunk - 1) / tiles_per_chunk;
  a.G = G; a.tiles_per_chunk = tiles_per_chunk; a.n_chunks = (int)n_chunks;
  const size_t need = (size_t)n_chunks * R * (sizeof(double) + 2 * sizeof(int32_t));
  int rc = ensure_workspace(p, need);
  if (rc) return rc;
  a.partL = reinterpret_cast<double*>(p->ws);
  a.partNc = reinterpret_cast<int32_t*>(a.partL + (size_t)n_chunks * R);
  a.partSq = a.partNc + (size_t)n_chunks * R;
  dim3 grid((unsigned)n_chunks, (unsigned)n_groups);
  const size_t smem = (size_t)G * per_chain;
  switch (p->DP) {
    case 8: rc = dispatch_op<8, 2>(p->OP, a, grid, smem, st); break;
    case 16: rc = dispatch_op<16, 2>(p->OP, a, grid, smem, st); break;
    case 32: rc = dispatch_op<32, 1>(p->OP, a, grid, smem, st); break;
    case 64: rc = dispatch_op<64, 1>(p->OP, a, grid, smem, st); break;
    default: set_error("fp32 path: D padded to %d unsupported (D <= 64)", p->DP); return SAPT_ERR_UNSUPPORTED;
  }
  if (rc) return rc;
  return launch_loglik_finalize(a.partL, a.partNc, a.partSq, (int)n_chunks, R, rows, n_rows, L, nc, sq, st);
}

int launch_loglik_finalize(const double* partL, const int32_t* partNc, const int32_t* partSq, int n_chunks, int R, const int32_t* rows,
                           const int32_t* n_rows, double* L, int32_t* nc, double* sq, cudaStream_t st) {
  const int wpb = 8;
  loglik_finalize_kernel<<<(R + wpb - 1) / wpb, wpb * 32, 0, st>>>(partL, partNc, partSq, n_chunks, R, rows, n_rows, L, nc, sq);
  SAPT_CHECK_LAUNCH();
  return SAPT_OK;
}

template <int DP>
static int fwd_dispatch(int OP, const LoglikF32Params& a, int64_t n, int32_t* fx, float* prob, size_t smem, cudaStream_t st) {
  const unsigned grid = (unsigned)((n + kThreads - 1) / kThreads);
#define SAPT_FWD(OPV)                                                                                   \
  case OPV: {                                                                                           \
    auto k = forward_rows_kernel<DP, OPV>;                                                              \
    if (smem > 48 * 1024) SAPT_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k<<<grid, kThreads, smem, st>>>(a, n, fx, prob);                                                    \
    SAPT_CHECK_LAUNCH();                                                                                \
    return SAPT_OK;                                                                                     \
  }
  switch (OP) {
    SAPT_FWD(4) SAPT_FWD(8) SAPT_FWD(12) SAPT_FWD(16) SAPT_FWD(20)
    default: set_error("O padded to %d unsupported", OP); return SAPT_ERR_UNSUPPORTED;
  }
#undef SAPT_FWD
}

int launch_forward_rows_f32(sapt_problem* p, int which, const float* theta, int32_t* fx, float* prob, cudaStream_t st) {
  const DataSet& ds = p->set[which];
  if (ds.n <= 0) return SAPT_OK;
  LoglikF32Params a{};
  a.x = ds.x; a.y = ds.y; a.n_pad = ds.n_pad; a.theta = theta; a.D = p->D; a.H = p->H; a.O = p->O; a.HP = p->HP;
  const size_t smem = (size_t)(p->DP * p->HP + p->HP * p->OP + p->HP + p->OP) * sizeof(float);
  switch (p->DP) {
    case 8: return fwd_dispatch<8>(p->OP, a, ds.n, fx, prob, smem, st);
    case 16: return fwd_dispatch<16>(p->OP, a, ds.n, fx, prob, smem, st);
    case 32: return fwd_dispatch<32>(p->OP, a, ds.n, fx, prob, smem, st);
    case 64: return fwd_dispatch<64>(p->OP, a, ds.n, fx, prob, smem, st);
    default: set_error("D padded to %d unsupported", p->DP); return SAPT_ERR_UNSUPPORTED;
  }
}

}  // namespace sapt
