// Per-chain MCMC kernels: proposal (+prior, +surrogate coin), Metropolis-Hastings accept with
// the reference's bookkeeping, the LG mid-step, and the sequential replica-exchange sweep.
// Reference lines are cited per kernel (RV=_revamp.py, RW/LG = the two misc_codeapril2020 forks).
#include <math.h>

#include "sapt_device.cuh"
#include "sapt_internal.h"
#include "sapt_mh.cuh"

namespace sapt {

// -------------------------------------------------------------------------------------------
// prior: -(P/2) ln s2 - sum(w^2)/(2 s2)                                        RV:404-410
// -------------------------------------------------------------------------------------------
__global__ void prior_kernel(const float* theta, int64_t ld, int P, double s2, double* prior) {
  __shared__ double scratch[32];
  const float* w = theta + (int64_t)blockIdx.x * ld;
  double acc = 0.0;
  for (int i = threadIdx.x; i < P; i += blockDim.x) { const double v = (double)w[i]; acc += v * v; }
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) prior[blockIdx.x] = prior_from_sumsq(acc, P, s2);
}

// -------------------------------------------------------------------------------------------
// proposal                                                       RV:521,525 / LG:551-560,651
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) propose_kernel(sapt_propose_args a) {
  __shared__ double scratch[32];
  __shared__ int s_sur;
  const int nrows = a.n_rows ? *a.n_rows : a.R;
  if ((int)blockIdx.x >= nrows) return;
  const int row = a.rows ? a.rows[blockIdx.x] : (int)blockIdx.x;
  const int slot = a.slot_of_row ? a.slot_of_row[row] : (row + a.slot_offset);
  const int step_i = a.step + (a.step_dev ? *a.step_dev : 0);
  if (threadIdx.x == 0) {
    int sur = 0;
    if (a.decide) {
      const double ku = a.ku_inject ? a.ku_inject[slot] : (double)philox_uniform(a.seed, SAPT_STREAM_KU, step_i, (uint32_t)slot, 0);
      sur = (a.surrogate_ready && ku < a.surrogate_prob && step_i >= a.gate + 1) ? 1 : 0;
      a.is_sur[row] = sur;
      if (sur) { const int k = atomicAdd(a.n_sur, 1); a.sur_rows[k] = row; }
      else     { const int k = atomicAdd(a.n_true, 1); a.true_rows[k] = row; }
    }
    s_sur = sur;
  }
  __syncthreads();
  const bool second = a.decide && a.redraw && s_sur;
  const float* zin = second ? a.z2_inject : a.z_inject;
  const int kind = second ? SAPT_STREAM_Z2 : a.kind;
  const float* __restrict__ b = a.base + (int64_t)row * a.ld;
  float* __restrict__ o = a.proposal + (int64_t)row * a.ld;
  const float step = (float)a.step_w;
  double acc = 0.0;
  if (zin) {
    const float* z = zin + (int64_t)slot * a.P;
    for (int i = threadIdx.x; i < a.P; i += blockDim.x) {
      const float v = fmaf(step, z[i], b[i]);
      o[i] = v;
      acc += (double)v * (double)v;
    }
  } else {
    const int nblk = (a.P + 3) >> 2;
    for (int q = threadIdx.x; q < nblk; q += blockDim.x) {
      float z[4];
      philox_normal4(a.seed, kind, step_i, (uint32_t)slot, (uint32_t)q, z);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = 4 * q + j;
        if (i < a.P) {
          const float v = fmaf(step, z[j], b[i]);
          o[i] = v;
          acc += (double)v * (double)v;
        }
      }
    }
  }
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) a.prior_prop[row] = prior_from_sumsq(acc, a.P, a.sigma_squared);
}

// -------------------------------------------------------------------------------------------
// LG: bookkeeping after the FIRST true evaluation + Langevin coin               LG:622-649
// -------------------------------------------------------------------------------------------
__global__ void lg_first_kernel(sapt_lg_first_args a) {
  const int n = a.n_true ? *a.n_true : a.R;
  if ((int)blockIdx.x >= n) return;
  const int row = a.true_rows ? a.true_rows[blockIdx.x] : (int)blockIdx.x;
  const int slot = a.st.slot_of_row[row];
  const int step_i = a.step + (a.step_dev ? *a.step_dev : 0);
  __shared__ int s_cnt;
  if (threadIdx.x == 0) {
    const double L = a.L_train[row];
    a.surg_first[row] = L / a.st.temperature[slot];
    int cnt = -1;
    if (a.train_count) {
      cnt = a.train_count[slot];
      if (cnt < a.train_cap) { a.train_L[(int64_t)slot * a.train_cap + cnt] = L; a.train_count[slot] = cnt + 1; }
      else cnt = -1;
    }
    s_cnt = cnt;
    if (a.use_langevin) {
      const double lx = a.lx2_inject ? a.lx2_inject[slot] : (double)philox_uniform(a.seed, SAPT_STREAM_LX2, step_i, (uint32_t)slot, 0);
      if (lx < a.l_prob) { const int k = atomicAdd(a.n_lang, 1); a.lang_rows[k] = row; }
    }
  }
  __syncthreads();
  if (s_cnt >= 0 && a.train_theta) {
    const float* src = a.proposal + (int64_t)row * a.ld;
    float* dst = a.train_theta + ((int64_t)slot * a.train_cap + s_cnt) * a.P;
    for (int i = threadIdx.x; i < a.P; i += blockDim.x) dst[i] = src[i];
  }
}

// -------------------------------------------------------------------------------------------
// Metropolis-Hastings accept + bookkeeping                                      RV:525-642
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) mh_accept_kernel(sapt_accept_args a) {
  const int row = blockIdx.x;
  const int slot = a.st.slot_of_row[row];
  const int S = a.hist.n_slots;
  const int step_i = a.step + (a.step_dev ? *a.step_dev : 0);
  __shared__ int s_accept, s_cnt;
  const int64_t hi = (int64_t)step_i * S + slot;        // history row i
  const int64_t ho = (int64_t)(step_i + 1) * S + slot;  // history row i+1
  if (threadIdx.x == 0) {
    MhInputs in;
    in.sur = a.is_sur && a.is_sur[row];
    in.sur_s = a.sur_s ? a.sur_s[row] : 0.f;
    in.L_train = a.L_train[row]; in.nc_train = a.nc_train[row]; in.sq_train = a.sq_train[row];
    in.L_test = a.n_test > 0 ? a.L_test[row] : 0.0; in.nc_test = a.n_test > 0 ? a.nc_test[row] : 0; in.sq_test = a.n_test > 0 ? a.sq_test[row] : 0.0;
    in.prior_prop = a.prior_prop[row];
    in.has_surg_first = a.surg_first != nullptr; in.surg_first = a.surg_first ? a.surg_first[row] : 0.0;
    in.ls_pred = a.ls_pred ? a.ls_pred[row] : 0.0; in.ls_avg = a.ls_avg ? a.ls_avg[row] : 0.0;
    int cnt;
    s_accept = mh_decide(a, slot, step_i, in, &cnt) ? 1 : 0;
    s_cnt = cnt;
  }
  __syncthreads();
  const bool acc = s_accept != 0;
  const float* __restrict__ prop = a.proposal + (int64_t)row * a.ld;
  float* __restrict__ th = a.theta + (int64_t)row * a.ld;
  float* tr = (s_cnt >= 0 && a.train_theta) ? a.train_theta + ((int64_t)slot * a.train_cap + s_cnt) * a.P : nullptr;
  float* pw_out = ((a.flags & SAPT_F_RECORD_POS) && a.hist.pos_w) ? a.hist.pos_w + ho * a.P : nullptr;
  const float* pw_in = pw_out ? a.hist.pos_w + hi * a.P : nullptr;
  for (int i = threadIdx.x; i < a.P; i += blockDim.x) {
    const float v = prop[i];
    if (tr) tr[i] = v;
    if (acc) th[i] = v;
    if (pw_out) pw_out[i] = acc ? v : pw_in[i];  // RV:590 / RV:608: reject copies HISTORY row i
  }
}

// -------------------------------------------------------------------------------------------
// replica exchange: sequential sweep over all chains                 RW:869-898, RW:1095-1099
// -------------------------------------------------------------------------------------------
__global__ void advance_kernel(int32_t* counter, int delta) { *counter += delta; }

// The sweep is a recurrence on ONE carried message: at pair k the carried message (what position k holds after
// the earlier pairs) competes with the message chain k+1 posted; the loser of the comparison stays at k, the
// other is carried on. Chains are staged through shared memory in chunks so any R fits; thread 0 walks a chunk,
// the CTA loads / scatters around it.
constexpr int kSweepChunk = 2048;

__global__ void swap_sweep_kernel(int R, double* lik_post, const double* temperature, int32_t* row_of_slot, uint32_t flags,
                                  uint32_t seed, int sweep_base, const int32_t* sweep_dev, const double* u_inject,
                                  uint8_t* swapped, int32_t* num_swap, int32_t* src_slot, int row0, int R_local,
                                  int32_t* slot_of_row_local, int32_t* scratch_src /*[R]*/, double* scratch_post /*[R]*/) {
  __shared__ double s_post[kSweepChunk];   // posted likelihood of chain k+1 (input) / of the message left at k (output)
  __shared__ double s_u[kSweepChunk];
  __shared__ int32_t s_src[kSweepChunk];   // output: which chain's message stays at position k
  __shared__ uint8_t s_sw[kSweepChunk];
  __shared__ double c_post;                // carried message
  __shared__ int c_src, c_ns;
  const int sweep = sweep_base + (sweep_dev ? *sweep_dev : 0);
  auto posted = [&](int k) {  // RW:623 posts likelihood*adapttemp ; LG:805 posts the tempered value
    return (flags & SAPT_F_POST_TEMPERED) ? lik_post[k] : lik_post[k] * temperature[k];
  };
  if (threadIdx.x == 0) { c_post = posted(0); c_src = 0; c_ns = 0; }
  for (int k0 = 0; k0 < R - 1; k0 += kSweepChunk) {
    const int n = min(kSweepChunk, R - 1 - k0);
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      const int k = k0 + j;
      s_post[j] = posted(k + 1);
      s_u[j] = u_inject ? u_inject[k] : (double)philox_uniform(seed, SAPT_STREAM_SWAP, sweep, 0u, (uint32_t)k);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double cp = c_post; int cs = c_src; int ns = c_ns;
      for (int j = 0; j < n; ++j) {
        const double np_ = s_post[j];
        const double d = fmin(709.0, np_ - cp);
        double pr = 0.5 * exp(d);
        if (d != d) pr = 1.0;  // min(1, nan) == 1 in Python
        pr = fmin(1.0, pr);
        const bool sw = s_u[j] < pr;
        if (sw) { s_src[j] = k0 + j + 1; s_post[j] = np_; ++ns; }       // chain k+1's message drops to position k
        else { s_src[j] = cs; s_post[j] = cp; cp = np_; cs = k0 + j + 1; } // carried message stays, k+1's is carried on
        s_sw[j] = sw ? 1 : 0;
      }
      c_post = cp; c_src = cs; c_ns = ns;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      const int k = k0 + j;
      scratch_src[k] = s_src[j];
      scratch_post[k] = s_post[j];
      if (swapped) swapped[k] = s_sw[j];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    scratch_src[R - 1] = c_src; scratch_post[R - 1] = c_post;
    if (num_swap) *num_swap += c_ns;
  }
  __syncthreads();
  // new binding: chain k now holds the vector chain src[k] posted. row_of_slot is permuted through a copy.
  int32_t* old_rows = scratch_src + R;  // [R]
  for (int k = threadIdx.x; k < R; k += blockDim.x) old_rows[k] = row_of_slot[k];
  __syncthreads();
  for (int k = threadIdx.x; k < R; k += blockDim.x) {
    const int newrow = old_rows[scratch_src[k]];
    row_of_slot[k] = newrow;
    if (src_slot) src_slot[k] = scratch_src[k];
    if (flags & SAPT_F_TAKE_POST_LIK) lik_post[k] = scratch_post[k];  // LG:813-815
    if (slot_of_row_local && newrow >= row0 && newrow < row0 + R_local) slot_of_row_local[newrow - row0] = k;
  }
}

__global__ void philox_fill_kernel(uint32_t seed, int kind, int step, int chain, int n, int normal, float* out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (normal) {
    if (4 * q >= n) return;
    float z[4];
    philox_normal4(seed, kind, step, (uint32_t)chain, (uint32_t)q, z);
    for (int j = 0; j < 4; ++j) if (4 * q + j < n) out[4 * q + j] = z[j];
  } else {
    if (q < n) out[q] = philox_uniform(seed, kind, step, (uint32_t)chain, (uint32_t)q);
  }
}

}  // namespace sapt

using namespace sapt;

extern "C" int sapt_prior(const float* theta, int64_t ld_theta, int R, int P, double sigma_squared, double* prior, void* stream) {
  SAPT_REQUIRE(theta && prior && R >= 0 && P > 0 && ld_theta >= P, "sapt_prior");
  if (R == 0) return SAPT_OK;
  prior_kernel<<<R, 256, 0, (cudaStream_t)stream>>>(theta, ld_theta, P, sigma_squared, prior);
  SAPT_CHECK_LAUNCH();
  return SAPT_OK;
}

extern "C" int sapt_propose(const sapt_propose_args* a, void* stream) {
  SAPT_REQUIRE(a && a->base && a->proposal && a->prior_prop && a->R >= 0 && a->P > 0 && a->ld >= a->P, "sapt_propose");
  if (a->decide) SAPT_REQUIRE(a->is_sur && a->true_rows && a->n_true && a->sur_rows && a->n_sur, "sapt_propose: decision outputs");
  if (a->R == 0) return SAPT_OK;
  propose_kernel<<<a->R, 256, 0, (cudaStream_t)stream>>>(*a);
  SAPT_CHECK_LAUNCH();
  return SAPT_OK;
}

extern "C" int sapt_lg_first_eval(const sapt_lg_first_args* a, void* stream) {
  SAPT_REQUIRE(a && a->proposal && a->L_train && a->surg_first && a->st.slot_of_row && a->st.temperature, "sapt_lg_first_eval");
  if (a->use_langevin) SAPT_REQUIRE(a->lang_rows && a->n_lang, "sapt_lg_first_eval: langevin outputs");
  if (a->R == 0) return SAPT_OK;
  lg_first_kernel<<<a->R, 128, 0, (cudaStream_t)stream>>>(*a);
  SAPT_CHECK_LAUNCH();
  return SAPT_OK;
}

extern "C" int sapt_mh_accept(const sapt_accept_args* a, void* stream) {
  SAPT_REQUIRE(a && a->theta && a->proposal && a->prior_prop && a->L_train, "sapt_mh_accept");
  SAPT_REQUIRE(a->st.slot_of_row && a->st.temperature && a->st.lik_cur && a->st.lik_post && a->st.prior_cur && a->st.mov &&
               a->st.likl_wo_temp && a->st.naccept && a->st.last_rmse_train && a->st.last_rmse_test &&
               a->st.last_acc_train && a->st.last_acc_test, "sapt_mh_accept: chain state");
  SAPT_REQUIRE(a->nc_train && a->sq_train && a->n_train > 0, "sapt_mh_accept: train stats");
  if (a->n_test > 0) SAPT_REQUIRE(a->L_test && a->nc_test && a->sq_test, "sapt_mh_accept: test stats");
  if (a->R == 0) return SAPT_OK;
  mh_accept_kernel<<<a->R, 128, 0, (cudaStream_t)stream>>>(*a);
  SAPT_CHECK_LAUNCH();
  return SAPT_OK;
}

extern "C" int sapt_advance(int32_t* counter, int delta, void* stream) {
  SAPT_REQUIRE(counter, "sapt_advance");
  advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter, delta);
  SAPT_CHECK_LAUNCH();
  return SAPT_OK;
}

extern "C" int sapt_swap_sweep(int R_global, double* lik_post, const double* temperature, int32_t* row_of_slot, uint32_t flags,
                               uint32_t seed, int sweep, const double* u_inject, uint8_t* swapped, int32_t* num_swap,
                               int32_t* src_slot, int row0, int R_local, int32_t* slot_of_row_local, void* stream) {
  return sapt_swap_sweep_dev(R_global, lik_post, temperature, row_of_slot, flags, seed, sweep, nullptr, u_inject, swapped, num_swap,
                             src_slot, row0, R_local, slot_of_row_local, stream);
}

extern "C" int sapt_swap_sweep_dev(int R_global, double* lik_post, const double* temperature, int32_t* row_of_slot, uint32_t flags,
                                   uint32_t seed, int sweep, const int32_t* sweep_dev, const double* u_inject, uint8_t* swapped,
                                   int32_t* num_swap, int32_t* src_slot, int row0, int R_local, int32_t* slot_of_row_local,
                                   void* stream) {
  SAPT_REQUIRE(R_global >= 1 && lik_post && temperature && row_of_slot, "sapt_swap_sweep");
  // scratch: src[R] + old_rows[R] (int32) and post[R] (double), one small buffer per device of this process
  constexpr int kMaxDev = 64;
  static int32_t* sc_i_dev[kMaxDev] = {nullptr}; static double* sc_d_dev[kMaxDev] = {nullptr}; static int sc_cap_dev[kMaxDev] = {0};
  int dev = 0;
  SAPT_CUDA(cudaGetDevice(&dev));
  SAPT_REQUIRE(dev >= 0 && dev < kMaxDev, "device ordinal");
  int32_t*& sc_i = sc_i_dev[dev]; double*& sc_d = sc_d_dev[dev]; int& sc_cap = sc_cap_dev[dev];
  if (R_global > sc_cap) {
    SAPT_CUDA(cudaDeviceSynchronize());
    if (sc_i) { SAPT_CUDA(cudaFree(sc_i)); SAPT_CUDA(cudaFree(sc_d)); }
    const int cap = R_global + R_global / 4 + 64;
    SAPT_CUDA(cudaMalloc(&sc_i, sizeof(int32_t) * 2 * cap));
    SAPT_CUDA(cudaMalloc(&sc_d, sizeof(double) * cap));
    sc_cap = cap;
  }
  // NOTE: old_rows lives at sc_i + R_global (the kernel derives it from scratch_src + R)
  swap_sweep_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(R_global, lik_post, temperature, row_of_slot, flags, seed, sweep, sweep_dev,
                                                         u_inject, swapped, num_swap, src_slot, row0, R_local, slot_of_row_local,
                                                         sc_i, sc_d);
  SAPT_CHECK_LAUNCH();
  return SAPT_OK;
}

extern "C" int sapt_philox_host(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  philox4x32_10(ctr, key[0], key[1], out);
  return SAPT_OK;
}

extern "C" int sapt_philox_fill(uint32_t seed, int kind, int step, int chain, int n, int normal, float* out, void* stream) {
  SAPT_REQUIRE(out && n >= 0, "sapt_philox_fill");
  if (n == 0) return SAPT_OK;
  const int work = normal ? (n + 3) / 4 : n;
  philox_fill_kernel<<<(work + 127) / 128, 128, 0, (cudaStream_t)stream>>>(seed, kind, step, chain, n, normal, out);
  SAPT_CHECK_LAUNCH();
  return SAPT_OK;
}
