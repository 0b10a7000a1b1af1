// Fused step kernel for SMALL problems (the reference's own runnable configurations: a few hundred rows, a handful
// of chains): the multi-kernel step is then bound by the latency of ~10 dependent launches (~67 us per step at C1),
// not by work. Here ONE launch runs n_steps complete MCMC steps of every chain: CTA = chain, theta lives in shared
// memory for the whole launch, each step does proposal (Philox) + prior, likelihood on train and test rows (rows
// strided over the CTA's threads, fp32 tiles of sapt_eval_f32.cuh), optional surrogate MLP, Metropolis-Hastings
// accept and bookkeeping (sapt_mh.cuh — the same code as mh_accept_kernel). Random-walk forks (rw / rv) only; the
// exchange sweep stays a separate launch between calls.
#include "sapt_device.cuh"
#include "sapt_eval_f32.cuh"
#include "sapt_internal.h"
#include "sapt_mh.cuh"

namespace sapt {

constexpr int kFT = 256;  // threads per CTA

struct FusedData {
  const float* x[2]; const int32_t* y[2]; int64_t n[2];
  int D, H, O, HP;
};

template <int DP, int OP>
__device__ __forceinline__ void fused_eval(const FusedData& d, int which, const float* sw, double* sc_d, int* sc_i, double& L, int& nc,
                                           double& sq) {
  double accL = 0.0; int accN = 0, accS = 0;
  const int64_t n = d.n[which];
  for (int64_t r = threadIdx.x; r < n; r += kFT) {
    float x[1][DP]; int y[1]; float lp[1]; int pred[1];
    const float4* xr = reinterpret_cast<const float4*>(d.x[which] + r * DP);
#pragma unroll
    for (int q = 0; q < DP / 4; ++q) { const float4 v = __ldg(xr + q); x[0][4 * q] = v.x; x[0][4 * q + 1] = v.y; x[0][4 * q + 2] = v.z; x[0][4 * q + 3] = v.w; }
    y[0] = __ldg(d.y[which] + r);
    eval_rows<DP, OP, 1>(sw, d.HP, d.O, x, y, lp, pred);
    accL += (double)lp[0];
    accN += (pred[0] == y[0]);
    const int e = pred[0] - y[0];
    accS += e * e;
  }
  L = block_sum(accL, sc_d);
  nc = block_sum(accN, sc_i);
  sq = (double)block_sum(accS, sc_i);
}

template <int DP, int OP>
__global__ void __launch_bounds__(kFT) fused_steps_kernel(FusedData d, sapt_fused_args a) {
  extern __shared__ __align__(16) float sm[];
  __shared__ double sc_d[32];
  __shared__ int sc_i[32];
  __shared__ int s_sur, s_accept, s_cnt;
  __shared__ float s_surout;
  const int row = blockIdx.x, tid = threadIdx.x;
  const int P = a.P;
  const int PS = (P + 3) & ~3;
  const int S = DP * d.HP + d.HP * OP + d.HP + OP;
  float* th = sm;               // [PS] current theta
  float* pr = th + PS;          // [PS] proposal
  float* sw = pr + PS;          // [S]  scaled weight image of the proposal
  float* a1 = sw + S;           // [h1]
  float* a2 = a1 + a.h1;        // [h2]
  const sapt_accept_args& m = a.acc;
  const int slot = m.st.slot_of_row[row];
  const int step0 = m.step + (m.step_dev ? *m.step_dev : 0);
  float* gth = a.theta + (int64_t)row * a.ld;
  for (int i = tid; i < P; i += kFT) th[i] = gth[i];
  __syncthreads();
  const float stepw = (float)a.step_w;
  for (int k = 0; k < a.n_steps; ++k) {
    const int step_i = step0 + k;
    if (tid == 0) {
      int sur = 0;
      if (a.use_surrogate) {
        const double ku = (double)philox_uniform(a.seed, SAPT_STREAM_KU, step_i, (uint32_t)slot, 0);
        sur = (a.surrogate_ready && ku < a.surrogate_prob && step_i >= a.gate + 1) ? 1 : 0;
      }
      s_sur = sur;
    }
    // ---- proposal + prior (RV:521, RV:404-410) ----
    double ss = 0.0;
    const int nblk = (P + 3) >> 2;
    for (int q = tid; q < nblk; q += kFT) {
      float z[4];
      philox_normal4(a.seed, SAPT_STREAM_Z1, step_i, (uint32_t)slot, (uint32_t)q, z);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = 4 * q + j;
        if (i < P) { const float v = fmaf(stepw, z[j], th[i]); pr[i] = v; ss += (double)v * (double)v; }
      }
    }
    ss = block_sum(ss, sc_d);  // (contains the barriers that publish pr[] and s_sur)
    const double prior_prop = prior_from_sumsq(ss, P, a.sigma_squared);
    const bool sur = s_sur != 0;
    MhInputs in;
    in.sur = sur; in.sur_s = 0.f; in.L_train = 0.0; in.nc_train = 0; in.sq_train = 0.0; in.L_test = 0.0; in.nc_test = 0; in.sq_test = 0.0;
    in.prior_prop = prior_prop; in.has_surg_first = false; in.surg_first = 0.0; in.ls_pred = 0.0; in.ls_avg = 0.0;
    const bool need_train = !sur || (m.flags & SAPT_F_COMPARE_SURROGATE);
    if (need_train) {
      stage_chain_weights<DP, OP>(pr, sw, d.D, d.H, d.O, d.HP, tid, kFT);
      __syncthreads();
      fused_eval<DP, OP>(d, 0, sw, sc_d, sc_i, in.L_train, in.nc_train, in.sq_train);
      if (!sur && d.n[1] > 0) fused_eval<DP, OP>(d, 1, sw, sc_d, sc_i, in.L_test, in.nc_test, in.sq_test);
    }
    if (sur) {
      // ---- surrogate MLP P -> h1 -> h2 -> 1 (relu, relu, sigmoid), fp32 like Keras (RV:217-233, 319) ----
      for (int j = tid; j < a.h1; j += kFT) {
        float acc = 0.f;
        for (int p = 0; p < P; ++p) acc = fmaf(pr[p], __ldg(a.W1 + (int64_t)p * a.h1 + j), acc);
        a1[j] = fmaxf(acc + a.b1[j], 0.f);
      }
      __syncthreads();
      for (int j = tid; j < a.h2; j += kFT) {
        float acc = 0.f;
        for (int q = 0; q < a.h1; ++q) acc = fmaf(a1[q], __ldg(a.W2 + q * a.h2 + j), acc);
        a2[j] = fmaxf(acc + a.b2[j], 0.f);
      }
      __syncthreads();
      if (tid < 32) {
        float acc = 0.f;
        for (int q = tid; q < a.h2; q += 32) acc = fmaf(a2[q], __ldg(a.W3 + q), acc);
        acc = warp_sum(acc);
        if (tid == 0) s_surout = 1.0f / (1.0f + expf(-(acc + a.b3[0])));
      }
      __syncthreads();
      in.sur_s = s_surout;
    }
    // ---- Metropolis-Hastings + bookkeeping (one thread; RV:525-642) ----
    if (tid == 0) {
      int cnt;
      s_accept = mh_decide(m, slot, step_i, in, &cnt) ? 1 : 0;
      s_cnt = cnt;
    }
    __syncthreads();
    const bool acc = s_accept != 0;
    const int64_t hi = (int64_t)step_i * m.hist.n_slots + slot, ho = hi + m.hist.n_slots;
    float* tr = (s_cnt >= 0 && m.train_theta) ? m.train_theta + ((int64_t)slot * m.train_cap + s_cnt) * P : nullptr;
    float* pw_out = ((m.flags & SAPT_F_RECORD_POS) && m.hist.pos_w) ? m.hist.pos_w + ho * P : nullptr;
    const float* pw_in = pw_out ? m.hist.pos_w + hi * P : nullptr;
    for (int i = tid; i < P; i += kFT) {
      const float v = pr[i];
      if (tr) tr[i] = v;
      if (pw_out) pw_out[i] = acc ? v : pw_in[i];
      if (acc) th[i] = v;
    }
    __syncthreads();
  }
  for (int i = tid; i < P; i += kFT) gth[i] = th[i];
}

}  // namespace sapt

using namespace sapt;

extern "C" int sapt_fused_steps(sapt_problem_t* p, const sapt_fused_args* a, void* stream) {
  SAPT_REQUIRE(p && a && a->theta && a->R >= 0 && a->P == p->P && a->ld >= a->P && a->n_steps >= 0, "sapt_fused_steps");
  const sapt_accept_args& m = a->acc;
  SAPT_REQUIRE(m.st.slot_of_row && m.st.temperature && m.st.lik_cur && m.st.lik_post && m.st.prior_cur && m.st.mov && m.st.likl_wo_temp &&
               m.st.naccept && m.st.last_rmse_train && m.st.last_rmse_test && m.st.last_acc_train && m.st.last_acc_test,
               "sapt_fused_steps: chain state");
  SAPT_REQUIRE(!(m.flags & SAPT_F_REDRAW_ON_SURROGATE) && m.surg_first == nullptr, "sapt_fused_steps: random-walk forks only");
  if (a->use_surrogate && a->surrogate_ready) SAPT_REQUIRE(a->W1 && a->b1 && a->W2 && a->b2 && a->W3 && a->b3 && a->h1 > 0 && a->h2 > 0, "surrogate weights");
  if (a->R == 0 || a->n_steps == 0) return SAPT_OK;
  FusedData d{};
  for (int w = 0; w < 2; ++w) { d.x[w] = p->set[w].x; d.y[w] = p->set[w].y; d.n[w] = p->set[w].n; }
  d.D = p->D; d.H = p->H; d.O = p->O; d.HP = p->HP;
  sapt_fused_args aa = *a;
  aa.acc.n_train = p->set[0].n; aa.acc.n_test = p->set[1].n; aa.acc.P = a->P; aa.acc.capture = 1;
  const int PS = (a->P + 3) & ~3;
  const size_t smem = (size_t)(2 * PS + p->DP * p->HP + p->HP * p->OP + p->HP + p->OP + std::max(a->h1, 0) + std::max(a->h2, 0)) * sizeof(float);
  if (smem > 200 * 1024) { set_error("sapt_fused_steps: topology needs %zu B of shared memory", smem); return SAPT_ERR_UNSUPPORTED; }
  cudaStream_t st = (cudaStream_t)stream;
#define SAPT_FUSED(DPV, OPV)                                                                                             \
  {                                                                                                                      \
    auto k = fused_steps_kernel<DPV, OPV>;                                                                               \
    if (smem > 48 * 1024) SAPT_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    k<<<a->R, kFT, smem, st>>>(d, aa);                                                                                   \
  }
#define SAPT_FUSED_OP(DPV)                                                                      \
  switch (p->OP) {                                                                              \
    case 4: SAPT_FUSED(DPV, 4) break;                                                           \
    case 8: SAPT_FUSED(DPV, 8) break;                                                           \
    case 12: SAPT_FUSED(DPV, 12) break;                                                         \
    case 16: SAPT_FUSED(DPV, 16) break;                                                         \
    case 20: SAPT_FUSED(DPV, 20) break;                                                         \
    default: set_error("sapt_fused_steps: O padded to %d unsupported", p->OP); return SAPT_ERR_UNSUPPORTED; \
  }
  switch (p->DP) {
    case 8: SAPT_FUSED_OP(8) break;
    case 16: SAPT_FUSED_OP(16) break;
    case 32: SAPT_FUSED_OP(32) break;
    case 64: SAPT_FUSED_OP(64) break;
    default: set_error("sapt_fused_steps: D padded to %d unsupported", p->DP); return SAPT_ERR_UNSUPPORTED;
  }
#undef SAPT_FUSED_OP
#undef SAPT_FUSED
  SAPT_CHECK_LAUNCH();
  return SAPT_OK;
}
