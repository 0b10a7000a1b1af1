// Library state, data-set upload and the likelihood entry points of the C-ABI (include/sapt.h).
#include <stdarg.h>
#include <string.h>

#include <vector>

#include "sapt_internal.h"

namespace sapt {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int ensure_workspace(sapt_problem* p, size_t bytes) {
  if (bytes <= p->ws_bytes) return SAPT_OK;
  // growing is a synchronising operation; it happens during warm-up, never in a steady-state step
  SAPT_CUDA(cudaDeviceSynchronize());
  if (p->ws) SAPT_CUDA(cudaFree(p->ws));
  p->ws = nullptr; p->ws_bytes = 0;
  const size_t want = bytes + bytes / 4;
  SAPT_CUDA(cudaMalloc(&p->ws, want));
  p->ws_bytes = want;
  return SAPT_OK;
}

static int pad_dim(int D) { return D <= 8 ? 8 : D <= 16 ? 16 : D <= 32 ? 32 : 64; }

static int upload_set(sapt_problem* p, int which, const float* x, const int32_t* y, int64_t n) {
  DataSet& ds = p->set[which];
  ds.n = n;
  ds.n_pad = ((n + 255) / 256) * 256;
  if (n == 0) return SAPT_OK;
  std::vector<float> xp((size_t)ds.n_pad * p->DP, 0.f);
  std::vector<int32_t> yp((size_t)ds.n_pad, -1);
  for (int64_t i = 0; i < n; ++i) {
    memcpy(&xp[(size_t)i * p->DP], x + (size_t)i * p->D, sizeof(float) * p->D);
    if (y[i] < 0 || y[i] >= p->O) { set_error("label %d at row %lld outside [0,%d)", y[i], (long long)i, p->O); return SAPT_ERR_INVALID; }
    yp[i] = y[i];
  }
  SAPT_CUDA(cudaMalloc(&ds.x, xp.size() * sizeof(float)));
  SAPT_CUDA(cudaMalloc(&ds.y, yp.size() * sizeof(int32_t)));
  SAPT_CUDA(cudaMemcpy(ds.x, xp.data(), xp.size() * sizeof(float), cudaMemcpyHostToDevice));
  SAPT_CUDA(cudaMemcpy(ds.y, yp.data(), yp.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  if (p->tc_ok) {
    int rc = tc_prepare_dataset(p, which, xp.data(), n);
    if (rc) return rc;
  }
  return SAPT_OK;
}

static int create_common(sapt_problem_t** out, int D, int H, int O, int device) {
  SAPT_REQUIRE(out, "out");
  SAPT_REQUIRE(D >= 1 && H >= 1 && O >= 1, "topology");
  if (D > 64 || O > 20 || H > 256) { set_error("topology %d-%d-%d outside the kernel envelope (D<=64, H<=256, O<=20)", D, H, O); return SAPT_ERR_UNSUPPORTED; }
  SAPT_CUDA(cudaSetDevice(device));
  sapt_problem* p = new sapt_problem();
  p->D = D; p->H = H; p->O = O; p->P = D * H + H * O + H + O;
  p->DP = pad_dim(D); p->HP = (H + 3) & ~3; p->OP = (O + 3) & ~3;
  p->device = device;
  cudaDeviceProp prop;
  SAPT_CUDA(cudaGetDeviceProperties(&prop, device));
  p->num_sms = prop.multiProcessorCount;
  p->tc_ok = (prop.major == 10) && tc_topology_supported(D, H, O);
  *out = p;
  return SAPT_OK;
}

}  // namespace sapt

using namespace sapt;

extern "C" int sapt_abi_version(void) { return SAPT_ABI_VERSION; }
extern "C" const char* sapt_last_error(void) { return g_err; }
extern "C" long long sapt_launch_count(void) { return g_launches.load(); }
extern "C" void sapt_launch_count_reset(void) { g_launches.store(0); }

extern "C" int sapt_problem_create_f32(sapt_problem_t** out, int D, int H, int O, const float* xtrain_host, const int32_t* ytrain_host,
                                       int64_t n_train, const float* xtest_host, const int32_t* ytest_host, int64_t n_test, int device) {
  SAPT_REQUIRE(n_train > 0 && xtrain_host && ytrain_host, "train set");
  SAPT_REQUIRE(n_test >= 0 && (n_test == 0 || (xtest_host && ytest_host)), "test set");
  int rc = create_common(out, D, H, O, device);
  if (rc) return rc;
  rc = upload_set(*out, 0, xtrain_host, ytrain_host, n_train);
  if (!rc) rc = upload_set(*out, 1, xtest_host, ytest_host, n_test);
  if (rc) { sapt_problem_destroy(*out); *out = nullptr; }
  return rc;
}

extern "C" int sapt_problem_create(sapt_problem_t** out, int D, int H, int O, const double* train_host, int64_t n_train,
                                   const double* test_host, int64_t n_test, int device) {
  SAPT_REQUIRE(n_train > 0 && train_host, "train set");
  SAPT_REQUIRE(n_test >= 0 && (n_test == 0 || test_host), "test set");
  SAPT_REQUIRE(D >= 1, "topology");
  auto split = [&](const double* src, int64_t n, std::vector<float>& x, std::vector<int32_t>& y) -> int {
    x.resize((size_t)n * D); y.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
      for (int d = 0; d < D; ++d) x[(size_t)i * D + d] = (float)src[(size_t)i * (D + 1) + d];
      const double lab = src[(size_t)i * (D + 1) + D];
      if (lab != (double)(int32_t)lab) { set_error("label %g at row %lld is not an integer class index", lab, (long long)i); return SAPT_ERR_INVALID; }
      y[i] = (int32_t)lab;
    }
    return SAPT_OK;
  };
  std::vector<float> xt, xs; std::vector<int32_t> yt, ys;
  int rc = split(train_host, n_train, xt, yt);
  if (!rc && n_test) rc = split(test_host, n_test, xs, ys);
  if (rc) return rc;
  return sapt_problem_create_f32(out, D, H, O, xt.data(), yt.data(), n_train, n_test ? xs.data() : nullptr, n_test ? ys.data() : nullptr, n_test, device);
}

extern "C" int sapt_problem_destroy(sapt_problem_t* p) {
  if (!p) return SAPT_OK;
  cudaSetDevice(p->device);
  for (int w = 0; w < 2; ++w) {
    if (p->set[w].x) cudaFree(p->set[w].x);
    if (p->set[w].y) cudaFree(p->set[w].y);
    if (p->set[w].a_bf16) cudaFree(p->set[w].a_bf16);
    if (p->set[w].tmap) free(p->set[w].tmap);
    if (p->set[w].a_f16) cudaFree(p->set[w].a_f16);
    if (p->set[w].tmap_f16) free(p->set[w].tmap_f16);
  }
  if (p->ws) cudaFree(p->ws);
  if (p->h2d_theta) cudaFree(p->h2d_theta);
  if (p->d_L) cudaFree(p->d_L);
  if (p->d_nc) cudaFree(p->d_nc);
  if (p->d_sq) cudaFree(p->d_sq);
  delete p;
  return SAPT_OK;
}

extern "C" int sapt_problem_num_param(const sapt_problem_t* p) { return p ? p->P : SAPT_ERR_INVALID; }
extern "C" int64_t sapt_problem_rows(const sapt_problem_t* p, int which) { return (p && (which == 0 || which == 1)) ? p->set[which].n : SAPT_ERR_INVALID; }
extern "C" int sapt_problem_tc_supported(const sapt_problem_t* p) { return (p && p->tc_ok) ? 1 : 0; }

namespace sapt { void tc2_work_model(int D, int H, int O, double* mufu, double* mma_flops); }
static constexpr int kTcAutoMinMacs = 128;   // SAPT_PATH_AUTO: D*H + H*O from which the tcgen05 likelihood kernel is used (see sapt_loglik)

// see sapt.h: the per-evaluation work the selected likelihood kernel EXECUTES
extern "C" int sapt_loglik_work_model(const sapt_problem_t* p, int path, double* out, int n_out) {
  SAPT_REQUIRE(p && out && n_out >= SAPT_WORK_COUNT, "problem / out[SAPT_WORK_COUNT]");
  if (path == SAPT_PATH_TC && !p->tc_ok) { set_error("tcgen05 path unavailable for topology %d-%d-%d", p->D, p->H, p->O); return SAPT_ERR_UNSUPPORTED; }
  const bool tc = (path == SAPT_PATH_TC) || (path == SAPT_PATH_AUTO && p->tc_ok && p->set[0].f16_range_ok && p->D * p->H + p->H * p->O >= kTcAutoMinMacs);
  const int D = p->D, H = p->H, O = p->O;
  out[SAPT_WORK_PATH] = tc ? SAPT_PATH_TC : SAPT_PATH_F32;
  out[SAPT_WORK_FLOPS_ALGORITHMIC] = 2.0 * D * H + 2.0 * H * O;
  if (tc) {
    sapt::tc2_work_model(D, H, O, &out[SAPT_WORK_MUFU], &out[SAPT_WORK_MMA_FLOPS]);
    out[SAPT_WORK_FFMA] = 0.0;
  } else {
    // fp32 tiles (sapt_eval_f32.cuh): hidden units padded to HP4 = 4*ceil(H/4), 2 MUFU per sigmoid, O output sigmoids + O exp + 1 lg2
    const int HP4 = p->HP;
    out[SAPT_WORK_MUFU] = 2.0 * HP4 + 3.0 * O + 1.0;
    out[SAPT_WORK_MMA_FLOPS] = 0.0;
    out[SAPT_WORK_FFMA] = (double)p->DP * HP4 + (double)HP4 * p->OP;
  }
  return SAPT_OK;
}

extern "C" int sapt_loglik(sapt_problem_t* p, int which, const float* theta, int64_t ld_theta, int R, const int32_t* rows,
                           const int32_t* n_rows, double* L, int32_t* ncorrect, double* sqerr, int path, void* stream) {
  SAPT_REQUIRE(p && (which == 0 || which == 1), "problem / set selector");
  SAPT_REQUIRE(theta && L && R >= 0 && ld_theta >= p->P, "theta / outputs");
  SAPT_REQUIRE((rows == nullptr) == (n_rows == nullptr), "rows and n_rows come together");
  if (path == SAPT_PATH_TC && !p->tc_ok) { set_error("tcgen05 path unavailable for topology %d-%d-%d on this device", p->D, p->H, p->O); return SAPT_ERR_UNSUPPORTED; }
  // AUTO: below kTcAutoMinMacs multiply-adds per evaluation the fp32 tiles win (measured crossover, devtools/tc_shapes.py, end of
  // round 2: 9-8-2 (88) and 4-12-3 (84) fp32 +20 %, 9-12-2 (132) tcgen05 +12 %, 8-16-4 (192) +6 %), above it the tensor-core path is
  // 1.1x-9x faster.
  const bool tc = (path == SAPT_PATH_TC) || (path == SAPT_PATH_AUTO && p->tc_ok && p->set[which].f16_range_ok && p->D * p->H + p->H * p->O >= kTcAutoMinMacs);
  if (tc) return launch_loglik_tc(p, which, theta, ld_theta, R, rows, n_rows, L, ncorrect, sqerr, (cudaStream_t)stream);
  return launch_loglik_f32(p, which, theta, ld_theta, R, rows, n_rows, L, ncorrect, sqerr, (cudaStream_t)stream);
}

extern "C" int sapt_loglik_host(sapt_problem_t* p, int which, const float* theta_host, int64_t ld_theta, int R, double* L_host,
                                int32_t* ncorrect_host, double* sqerr_host, int path) {
  SAPT_REQUIRE(p && theta_host && L_host && R > 0 && ld_theta >= p->P, "sapt_loglik_host");
  const size_t tb = (size_t)R * ld_theta * sizeof(float);
  if (tb > p->h2d_theta_bytes) {
    if (p->h2d_theta) SAPT_CUDA(cudaFree(p->h2d_theta));
    p->h2d_theta = nullptr; p->h2d_theta_bytes = 0;
    SAPT_CUDA(cudaMalloc(&p->h2d_theta, tb));
    p->h2d_theta_bytes = tb;
  }
  if (R > p->out_cap) {
    if (p->d_L) { SAPT_CUDA(cudaFree(p->d_L)); SAPT_CUDA(cudaFree(p->d_nc)); SAPT_CUDA(cudaFree(p->d_sq)); }
    SAPT_CUDA(cudaMalloc(&p->d_L, R * sizeof(double)));
    SAPT_CUDA(cudaMalloc(&p->d_nc, R * sizeof(int32_t)));
    SAPT_CUDA(cudaMalloc(&p->d_sq, R * sizeof(double)));
    p->out_cap = R;
  }
  cudaStream_t st = 0;
  SAPT_CUDA(cudaMemcpyAsync(p->h2d_theta, theta_host, tb, cudaMemcpyHostToDevice, st));
  int rc = sapt_loglik(p, which, p->h2d_theta, ld_theta, R, nullptr, nullptr, p->d_L, p->d_nc, p->d_sq, path, st);
  if (rc) return rc;
  SAPT_CUDA(cudaMemcpyAsync(L_host, p->d_L, R * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (ncorrect_host) SAPT_CUDA(cudaMemcpyAsync(ncorrect_host, p->d_nc, R * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (sqerr_host) SAPT_CUDA(cudaMemcpyAsync(sqerr_host, p->d_sq, R * sizeof(double), cudaMemcpyDeviceToHost, st));
  SAPT_CUDA(cudaStreamSynchronize(st));
  return SAPT_OK;
}

extern "C" int sapt_evaluate_proposal(sapt_problem_t* p, int which, const float* theta, int32_t* fx, float* prob, void* stream) {
  SAPT_REQUIRE(p && (which == 0 || which == 1) && theta && fx && prob, "sapt_evaluate_proposal");
  return launch_forward_rows_f32(p, which, theta, fx, prob, (cudaStream_t)stream);
}

// layout self-check for FFI bindings: sizes of the argument structs, in header order
extern "C" int sapt_struct_sizes(int32_t out[8]) {
  out[0] = (int32_t)sizeof(sapt_propose_args);
  out[1] = (int32_t)sizeof(sapt_chain_state);
  out[2] = (int32_t)sizeof(sapt_history);
  out[3] = (int32_t)sizeof(sapt_accept_args);
  out[4] = (int32_t)sizeof(sapt_lg_first_args);
  out[5] = (int32_t)sizeof(sapt_fused_args);
  out[6] = (int32_t)sizeof(sapt_local_surrogate);
  out[7] = (int32_t)sizeof(sapt_step_args);
  return SAPT_OK;
}

// n MCMC steps of the random-walk forks without host code between the kernels (the order is PTEngine.step's)
extern "C" int sapt_step(sapt_problem_t* p, const sapt_step_args* a, void* stream) {
  SAPT_REQUIRE(p && a && a->n_steps >= 0 && a->counters && a->L_train && a->nc_train && a->sq_train, "sapt_step");
  SAPT_REQUIRE(!a->prop.step_dev && !a->acc.step_dev, "sapt_step advances the step index itself: step_dev must be NULL");
  sapt_propose_args prop = a->prop;
  sapt_accept_args acc = a->acc;
  const int R = prop.R;
  const int64_t ntest = p->set[1].n;
  cudaStream_t st = (cudaStream_t)stream;
  for (int s = 0; s < a->n_steps; ++s) {
    const int step = a->prop.step + s;
    SAPT_CUDA(cudaMemsetAsync(a->counters, 0, 3 * sizeof(int32_t), st));
    prop.step = step;
    int rc = sapt_propose(&prop, stream);
    if (rc) return rc;
    const bool use_sur = a->use_surrogate != 0;
    const bool sur_possible = use_sur && prop.surrogate_ready && prop.surrogate_prob > 0 && step >= prop.gate + 1;
    if (sur_possible) {
      SAPT_REQUIRE(a->W1 && a->sur_s, "sapt_step: surrogate weights");
      rc = sapt_surrogate_predict(prop.proposal, prop.ld, R, prop.P, prop.sur_rows, prop.n_sur, a->W1, a->b1, a->h1, a->W2, a->b2, a->h2,
                                  a->W3, a->b3, a->sur_s, stream);
      if (rc) return rc;
    }
    const int32_t* true_rows = use_sur ? prop.true_rows : nullptr;
    const int32_t* n_true = use_sur ? prop.n_true : nullptr;
    const bool all_rows = a->compare_surrogate || !use_sur;                         // RV:543-544
    rc = sapt_loglik(p, 0, prop.proposal, prop.ld, R, all_rows ? nullptr : true_rows, all_rows ? nullptr : n_true, a->L_train,
                     a->nc_train, a->sq_train, a->path, stream);
    if (rc) return rc;
    if (ntest > 0) {
      rc = sapt_loglik(p, 1, prop.proposal, prop.ld, R, true_rows, n_true, a->L_test, a->nc_test, a->sq_test, a->path, stream);
      if (rc) return rc;
    }
    acc.step = step;
    rc = sapt_mh_accept(&acc, stream);
    if (rc) return rc;
  }
  return SAPT_OK;
}
