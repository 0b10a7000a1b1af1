/* A host WITHOUT Python or PyTorch: plain C over the C-ABI of include/sapt.h and the CUDA runtime.
 * Runs random-walk parallel tempering (the loop of ptReplica.run + the coordinator's exchange sweep,
 * _revamp.py:505-642 / surrogate_rw_common_interval.py:1095-1099) for R chains on synthetic data and prints the
 * final state as one JSON line. tests/test_gpu_parity.py runs this binary and checks that PTEngine, given the
 * same data, temperatures and seed, ends in the same state (same kernels, same Philox streams).
 *
 *   build: make -C examples/c_host        (gcc, links ../../sapt_b200/libsapt.so)
 *   usage: run_pt R samples swap_interval seed
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sapt.h"

#define CK(x) do { int rc_ = (x); if (rc_) { fprintf(stderr, "%s failed (%d): %s\n", #x, rc_, sapt_last_error()); exit(1); } } while (0)
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

static void* dalloc(size_t bytes) { void* p; CU(cudaMalloc(&p, bytes)); CU(cudaMemset(p, 0, bytes)); return p; }

int main(int argc, char** argv) {
  const int R = argc > 1 ? atoi(argv[1]) : 8;
  const int samples = argc > 2 ? atoi(argv[2]) : 40;
  const int swap_interval = argc > 3 ? atoi(argv[3]) : 5;
  const unsigned seed = argc > 4 ? (unsigned)atoi(argv[4]) : 7u;
  enum { D = 9, H = 12, O = 2, NTR = 300, NTE = 120 };
  const int P = D * H + H * O + H + O, ld = (P + 3) & ~3;

  /* synthetic data by formula (the Python test rebuilds the same arrays) */
  float* xtr = malloc(sizeof(float) * NTR * D); int* ytr = malloc(sizeof(int) * NTR);
  float* xte = malloc(sizeof(float) * NTE * D); int* yte = malloc(sizeof(int) * NTE);
  for (int i = 0; i < NTR; ++i) { for (int d = 0; d < D; ++d) xtr[i * D + d] = (float)(((i * 37 + d * 11) % 23) - 11) / 8.0f; ytr[i] = (i * 7 + 3) % O; }
  for (int i = 0; i < NTE; ++i) { for (int d = 0; d < D; ++d) xte[i * D + d] = (float)(((i * 29 + d * 13) % 19) - 9) / 6.0f; yte[i] = (i * 5 + 1) % O; }
  sapt_problem_t* prob;
  CK(sapt_problem_create_f32(&prob, D, H, O, xtr, ytr, NTR, xte, yte, NTE, 0));

  /* device state: theta rows, chain scalars indexed by chain id, the row <-> chain binding */
  float* theta = dalloc(sizeof(float) * R * ld); float* prop = dalloc(sizeof(float) * R * ld);
  double *T = dalloc(8 * R), *lik_cur = dalloc(8 * R), *lik_post = dalloc(8 * R), *prior_cur = dalloc(8 * R), *mov = dalloc(8 * 3 * R);
  double *lwt = dalloc(8 * R), *last = dalloc(8 * 4 * R), *prior_prop = dalloc(8 * R);
  double *L_tr = dalloc(8 * R), *sq_tr = dalloc(8 * R), *L_te = dalloc(8 * R), *sq_te = dalloc(8 * R);
  int32_t *nc_tr = dalloc(4 * R), *nc_te = dalloc(4 * R), *naccept = dalloc(4 * R), *slot_of_row = dalloc(4 * R), *row_of_slot = dalloc(4 * R);
  int32_t *is_sur = dalloc(4 * R), *true_rows = dalloc(4 * R), *sur_rows = dalloc(4 * R), *counters = dalloc(4 * 4);
  float* sur_s = dalloc(4 * R);

  /* all chains start from the same w (RV:940-943); temperatures: geometric ladder 1 .. 4 */
  float* h_theta = calloc((size_t)R * ld, sizeof(float));
  double* h_T = malloc(8 * R); int32_t* h_id = malloc(4 * R);
  for (int r = 0; r < R; ++r) {
    for (int k = 0; k < P; ++k) h_theta[(size_t)r * ld + k] = (float)(((k * 17 + 5) % 41) - 20) / 40.0f;
    h_T[r] = R > 1 ? pow(4.0, (double)r / (double)(R - 1)) : 1.0;
    h_id[r] = r;
  }
  CU(cudaMemcpy(theta, h_theta, sizeof(float) * R * ld, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(T, h_T, 8 * R, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(slot_of_row, h_id, 4 * R, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(row_of_slot, h_id, 4 * R, cudaMemcpyHostToDevice));

  /* initial likelihood / prior (RV:452-468) */
  double *h_L = malloc(8 * R), *h_Lte = malloc(8 * R), *h_pr = malloc(8 * R);
  CK(sapt_loglik(prob, 0, theta, ld, R, NULL, NULL, L_tr, nc_tr, sq_tr, SAPT_PATH_AUTO, NULL));
  CK(sapt_loglik(prob, 1, theta, ld, R, NULL, NULL, L_te, nc_te, sq_te, SAPT_PATH_AUTO, NULL));
  CK(sapt_prior(theta, ld, R, P, 25.0, prior_prop, NULL));
  CU(cudaMemcpy(h_L, L_tr, 8 * R, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(h_Lte, L_te, 8 * R, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(h_pr, prior_prop, 8 * R, cudaMemcpyDeviceToHost));
  for (int r = 0; r < R; ++r) h_L[r] /= h_T[r];
  CU(cudaMemcpy(lik_cur, h_L, 8 * R, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(lik_post, h_L, 8 * R, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(prior_cur, h_pr, 8 * R, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(lwt, h_Lte, 8 * R, cudaMemcpyHostToDevice));   /* RW:466-467: the TEST value survives */

  sapt_step_args a;
  memset(&a, 0, sizeof a);
  a.prop.base = theta; a.prop.proposal = prop; a.prop.ld = ld; a.prop.R = R; a.prop.P = P;
  a.prop.slot_of_row = slot_of_row; a.prop.step_w = 0.025; a.prop.sigma_squared = 25.0; a.prop.prior_prop = prior_prop;
  a.prop.seed = seed; a.prop.kind = SAPT_STREAM_Z1;
  a.prop.is_sur = is_sur; a.prop.true_rows = true_rows; a.prop.n_true = counters; a.prop.sur_rows = sur_rows; a.prop.n_sur = counters + 1;
  a.acc.theta = theta; a.acc.proposal = prop; a.acc.ld = ld; a.acc.R = R; a.acc.P = P;
  a.acc.st.slot_of_row = slot_of_row; a.acc.st.temperature = T; a.acc.st.lik_cur = lik_cur; a.acc.st.lik_post = lik_post;
  a.acc.st.prior_cur = prior_cur; a.acc.st.mov = mov; a.acc.st.likl_wo_temp = lwt;
  a.acc.st.last_rmse_train = last; a.acc.st.last_rmse_test = last + R; a.acc.st.last_acc_train = last + 2 * R; a.acc.st.last_acc_test = last + 3 * R;
  a.acc.st.naccept = naccept;
  a.acc.hist.n_slots = R;                                   /* no history kept */
  a.acc.L_train = L_tr; a.acc.nc_train = nc_tr; a.acc.sq_train = sq_tr; a.acc.L_test = L_te; a.acc.nc_test = nc_te; a.acc.sq_test = sq_te;
  a.acc.n_train = NTR; a.acc.n_test = NTE; a.acc.prior_prop = prior_prop; a.acc.sur_s = sur_s; a.acc.seed = seed; a.acc.capture = 1;
  a.sur_s = sur_s; a.L_train = L_tr; a.nc_train = nc_tr; a.sq_train = sq_tr; a.L_test = L_te; a.nc_test = nc_te; a.sq_test = sq_te;
  a.counters = counters; a.path = SAPT_PATH_AUTO;

  /* the chain loop: stretches of steps between two exchanges go to sapt_step in one call */
  int32_t* num_swap = counters + 3;
  const int n = samples - 1;
  int sweep = 0;
  for (int i = 0; i < n;) {
    int j = i;
    while (!(j % swap_interval == 0 && j != 0) && j < n - 1) ++j;
    a.prop.step = i; a.acc.step = i; a.n_steps = j - i + 1;
    CK(sapt_step(prob, &a, NULL));
    if (j % swap_interval == 0 && j != 0 && R > 1)
      CK(sapt_swap_sweep(R, lik_post, T, row_of_slot, 0, seed, sweep++, NULL, NULL, num_swap, NULL, 0, R, slot_of_row, NULL));
    i = j + 1;
  }
  CU(cudaDeviceSynchronize());

  int32_t* h_acc = malloc(4 * R); int32_t* h_ros = malloc(4 * R); int32_t h_cnt[4];
  CU(cudaMemcpy(h_acc, naccept, 4 * R, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(h_ros, row_of_slot, 4 * R, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(h_L, lik_cur, 8 * R, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(h_cnt, counters, 16, cudaMemcpyDeviceToHost));
  printf("{\"R\": %d, \"samples\": %d, \"num_swap\": %d, \"launches\": %lld, \"naccept\": [", R, samples, h_cnt[3], sapt_launch_count());
  for (int r = 0; r < R; ++r) printf("%d%s", h_acc[r], r + 1 < R ? ", " : "");
  printf("], \"row_of_slot\": [");
  for (int r = 0; r < R; ++r) printf("%d%s", h_ros[r], r + 1 < R ? ", " : "");
  printf("], \"lik_cur\": [");
  for (int r = 0; r < R; ++r) printf("%.17g%s", h_L[r], r + 1 < R ? ", " : "");
  printf("]}\n");
  CK(sapt_problem_destroy(prob));
  return 0;
}
