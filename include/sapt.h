/*
 * sapt.h — C-ABI of the B200-native surrogate-assisted parallel-tempering hot path.
 *
 * The reference (sydney-machine-learning/surrogate-assisted-parallel-tempering) has no
 * plugin / FFI layer: its boundary is the Python class surface. Each entry
 * point below names the reference function(s) it replaces, batched over the R replicas
 * (chains) a GPU hosts. File abbreviations: RV=_revamp.py,
 * RW=misc_codeapril2020/surrogate_rw_common_interval.py,
 * LG=misc_codeapril2020/surrogate_pt_classifier_langevingrad.py.
 *
 * Conventions
 *   - every function returns 0 on success, a negative SAPT_ERR_* otherwise; no exception
 *     crosses the boundary; sapt_last_error() gives a thread-local message.
 *   - pointers are DEVICE pointers unless the name ends in _host.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream). Nothing here
 *     synchronises the host except sapt_problem_create / *_host entry points.
 *   - theta is fp32 [R][ld_theta], one chain per row, element order of Network.encode
 *     (RV:120-127): W1 row-major [D][H], W2 row-major [H][O], B1[H], B2[O].
 *   - "slot" = chain identity (temperature ladder position, RNG stream, history row);
 *     "row" = where a parameter vector lives. Replica exchange re-binds slots to rows
 *     instead of moving parameter vectors.
 */
#ifndef SAPT_H_
#define SAPT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAPT_ABI_VERSION 1

enum {
  SAPT_OK = 0,
  SAPT_ERR_INVALID = -1,     /* bad argument (NULL, negative size, label out of range ...) */
  SAPT_ERR_UNSUPPORTED = -2, /* topology outside the compiled kernel envelope */
  SAPT_ERR_CUDA = -3,        /* a CUDA runtime call failed; see sapt_last_error() */
  SAPT_ERR_NOMEM = -4
};

/* likelihood kernel selection */
enum {
  SAPT_PATH_AUTO = 0, /* tcgen05 path when the topology fits it, else fp32 */
  SAPT_PATH_F32 = 1,  /* fp32 CUDA-core tiles */
  SAPT_PATH_TC = 2    /* tcgen05/TMEM stacked GEMM (bf16x3 split, fp32 accumulate) */
};

/* random stream kinds (counter word 2 of the Philox4x32-10 block); oracle/sapt_oracle.py KIND_* */
enum {
  SAPT_STREAM_Z1 = 0,  /* first proposal normals           np.random.normal  LG:551 */
  SAPT_STREAM_Z2 = 1,  /* redraw / Langevin proposal noise np.random.normal  LG:560,651 */
  SAPT_STREAM_LX1 = 2, /* unused uniform                   np.random.uniform LG:530 */
  SAPT_STREAM_KU = 3,  /* surrogate coin                   random.uniform    LG:553 */
  SAPT_STREAM_LX2 = 4, /* Langevin coin                    np.random.uniform LG:647 */
  SAPT_STREAM_U = 5,   /* Metropolis-Hastings coin         random.uniform    LG:704 */
  SAPT_STREAM_SWAP = 6, /* exchange coin                   np.random.uniform RW:888 */
  SAPT_STREAM_LS_INIT = 7 /* local-surrogate net init        np.random.randn   LS:62-65 */
};

/* behaviour switches of the reference's forks (SURVEY.md §2 "variant differences") */
enum {
  SAPT_F_POST_TEMPERED = 1 << 0,   /* exchange message carries tempered L (LG:805), else L*T (RW:623) */
  SAPT_F_TAKE_POST_LIK = 1 << 1,   /* `likelihood = result[P+1]` after the exchange (LG:813-815) */
  SAPT_F_DENORM_HALF_MIN = 1 << 2, /* surrogate de-normalisation s*minY/2 (LG:325), else s*(maxY-minY)+minY (RV:320) */
  SAPT_F_COMPARE_SURROGATE = 1 << 3, /* also evaluate the true train likelihood on surrogate steps (RV:543-544) */
  SAPT_F_REDRAW_ON_SURROGATE = 1 << 4, /* surrogate branch redraws the proposal from stream Z2 (LG:560) */
  SAPT_F_RECORD_POS = 1 << 5,      /* keep the full pos_w history on the device */
  SAPT_F_LOCAL_SURROGATE = 1 << 6, /* surrogate steps take sapt_local_surrogate_predict's value as is (LS:598,615-617) */
  SAPT_F_SWEEP_PARALLEL = 1 << 7,  /* sapt_swap_sweep: use the parallel (sparse table + pointer doubling) kernel even for a short ladder
                                      (default: R_global >= 4096); same decisions as the serial walk */
  SAPT_F_SWEEP_GRID = 1 << 8       /* with SAPT_F_SWEEP_PARALLEL: run its phases as grid-wide launches even for a short ladder (default:
                                      R_global >= 16384; the one-CTA kernel is latency bound: 1.6 ms at 65 536 chains); same decisions */
};

typedef struct sapt_problem sapt_problem_t; /* opaque: topology + device-resident data sets */

/* ---- library ------------------------------------------------------------------------ */
int sapt_abi_version(void);
const char* sapt_last_error(void);
/* number of kernels this library has launched since load / since the last reset
 * (bench.py's "gpu_launches" is read from here, not guessed) */
long long sapt_launch_count(void);
void sapt_launch_count_reset(void);
/* sizeof() of sapt_propose_args, sapt_chain_state, sapt_history, sapt_accept_args, sapt_lg_first_args,
 * sapt_fused_args, sapt_local_surrogate, sapt_step_args — lets an FFI binding verify its struct mirrors at load time */
int sapt_struct_sizes(int32_t out[8]);

/* ---- data + topology: replaces Network.__init__(Topo, Train, Test, lr) RV:49-60 -------
 * train_host/test_host: row-major float64 [n][D+1], features then the class index as a
 * float in column D — exactly the reference's `traindata` / `testdata` arrays (RV:1425-1519).
 * Copies to HBM once (fp32 features padded to a 16-byte multiple, int32 labels, and the
 * bf16-split operand image for the tcgen05 path). Labels must be integers in [0, O). */
int sapt_problem_create(sapt_problem_t** out, int D, int H, int O,
                        const double* train_host, int64_t n_train,
                        const double* test_host, int64_t n_test, int device);
/* same, from float32 features [n][D] and int32 labels (synthetic benchmark shapes) */
int sapt_problem_create_f32(sapt_problem_t** out, int D, int H, int O,
                            const float* xtrain_host, const int32_t* ytrain_host, int64_t n_train,
                            const float* xtest_host, const int32_t* ytest_host, int64_t n_test, int device);
int sapt_problem_destroy(sapt_problem_t* p);
int sapt_problem_num_param(const sapt_problem_t* p); /* P = D*H + H*O + H + O */
int64_t sapt_problem_rows(const sapt_problem_t* p, int which);
/* 1 if the tcgen05 path supports this topology */
int sapt_problem_tc_supported(const sapt_problem_t* p);

/* ---- batched likelihood: replaces ptReplica.likelihood_func RV:391-402 -> Network.evaluate_proposal
 * RV:155-174 -> ForwardPass RV:70-76 + softmax RV:129-131, for R chains at once ------------
 * which: 0 = train set, 1 = test set.
 * rows/n_rows (device, nullable): evaluate only theta rows rows[0..*n_rows); others untouched.
 * Outputs, indexed by theta row: L = sum_i log softmax(out_i)[y_i] (UNtempered, fp64
 * accumulation), ncorrect = #{argmax(out_i) == y_i} (first max on ties, like np.argmax),
 * sqerr = sum_i (argmax_i - y_i)^2 (rmse^2 * n of RV:380-382). */
int sapt_loglik(sapt_problem_t* p, int which, const float* theta, int64_t ld_theta, int R,
                const int32_t* rows, const int32_t* n_rows,
                double* L, int32_t* ncorrect, double* sqerr, int path, void* stream);
/* host-buffer convenience (the e2e path): H2D of theta, kernel, D2H of the three outputs */
int sapt_loglik_host(sapt_problem_t* p, int which, const float* theta_host, int64_t ld_theta, int R,
                     double* L_host, int32_t* ncorrect_host, double* sqerr_host, int path);
/* per-row outputs of Network.evaluate_proposal for ONE theta (RV:155-174):
 * fx[n] = argmax class, prob[n][O] = softmax over the sigmoid outputs */
int sapt_evaluate_proposal(sapt_problem_t* p, int which, const float* theta,
                           int32_t* fx, float* prob, void* stream);

/* ---- prior: replaces ptReplica.prior_likelihood RV:404-410 ----------------------------- */
int sapt_prior(const float* theta, int64_t ld_theta, int R, int P, double sigma_squared,
               double* prior, void* stream);

/* ---- proposal: replaces `np.random.normal(w, step_w, w_size)` RV:521 / LG:551,560,651 and
 * the surrogate coin `ku < surrogate_prob and i >= interval + 1` RV:525 -----------------------
 * For each listed row r (rows/n_rows nullable = all R): slot s = slot_of_row[r];
 *   proposal[r] = base[r] + step_w * z,  z = z_inject[s][:] if z_inject else Philox(seed, kind, step, s)
 *   prior_prop[r] = prior_likelihood(sigma_squared, proposal[r])
 * If decide != 0 the surrogate coin is also thrown (ku_inject[s] or Philox KU):
 *   is_sur[r] = surrogate_ready && ku < surrogate_prob && step >= gate + 1
 *   rows with is_sur==0 are appended to true_rows (count *n_true), the others to sur_rows.
 *   With redraw != 0 surrogate rows use z2_inject / stream Z2 instead (LG:560). */
typedef struct {
  const float* base;        /* [R][ld] current theta (or theta_gd for the Langevin proposal) */
  float* proposal;          /* [R][ld] */
  int64_t ld;
  int R, P;
  const int32_t* slot_of_row; /* [R] (nullable = identity) */
  int slot_offset;          /* added to the local slot id to form the global chain id */
  const int32_t* rows;      /* nullable */
  const int32_t* n_rows;    /* nullable */
  double step_w, sigma_squared;
  double* prior_prop;       /* [R] */
  uint32_t seed; int kind; int step;
  const float* z_inject;    /* [R_slots][P] or NULL */
  /* surrogate decision */
  int decide; int redraw; int surrogate_ready; double surrogate_prob; int gate;
  const float* z2_inject;   /* [R_slots][P] or NULL */
  const double* ku_inject;  /* [R_slots] or NULL */
  int32_t* is_sur;          /* [R] */
  int32_t* true_rows; int32_t* n_true; int32_t* sur_rows; int32_t* n_sur;
  /* CUDA-graph replay: if non-NULL the effective step is step + *step_dev (read on the device), so a captured
   * sequence of launches can be replayed for later steps without re-recording (see sapt_advance) */
  const int32_t* step_dev;
} sapt_propose_args;
int sapt_propose(const sapt_propose_args* a, void* stream);

/* ---- Langevin "gradient": replaces Network.langevin_gradient LG:138-156 + BackwardPass
 * LG:96-106 (one epoch of per-row online SGD, sequential over the train rows) for the listed
 * rows. mode 0 = "sgd_exact" (reference semantics); mode 1 = "batch": theta + lr * scale *
 * sum_i g_i(theta) with g_i the LG:100-106 increments at FIXED theta — on the tensor cores
 * (tcgen05: forward, back-propagation and the X^T.delta / h^T.delta row contractions) where the topology allows
 * (O <= 16, H <= 64, D <= 64), else on CUDA cores; mode 2 = "batch" on CUDA cores always (cross-check).
 * SAPT_ERR_UNSUPPORTED when the topology exceeds the kernels' shared-memory envelope (H beyond ~150). */
int sapt_langevin(sapt_problem_t* p, const float* theta, float* theta_gd, int64_t ld, int R,
                  const int32_t* rows, const int32_t* n_rows, double lrate, int mode,
                  double batch_scale, void* stream);

/* ---- surrogate predict: replaces surrogate.predict RV:297-322 (Keras MLP P->h1->h2->1,
 * relu/relu/sigmoid, RV:217-233) for the listed rows; s[r] = raw network output in (0,1).
 * Weights in Keras Dense layout: W1[P][h1], b1[h1], W2[h1][h2], b2[h2], W3[h2], b3[1]. */
int sapt_surrogate_predict(const float* theta, int64_t ld, int R, int P,
                           const int32_t* rows, const int32_t* n_rows,
                           const float* W1, const float* b1, int h1,
                           const float* W2, const float* b2, int h2,
                           const float* W3, const float* b3, float* s, void* stream);

/* ---- surrogate rendezvous: the interval's training rows (theta', L_untempered) that sapt_mh_accept / sapt_lg_first_eval
 * appended to the slot-indexed buffers train_theta[n_slots][cap][P], train_L[n_slots][cap] — replaces the (n, P+1) float64
 * blocks the chains put on surrogate_parameterqueue (RW:626-627, LG:822) and their concatenation chain by chain
 * (RW:1105-1116). Index vectors come from the caller (static shapes, computed on the device from the per-chain counts).
 * pack: payload[i] = [theta' (P fp32) | L (fp64 bits, two words) | dest[i] | 0], row width P + 4 words, for
 *       src[i] = slot * cap + j >= 0; src[i] < 0 marks a padding row (destination -1).
 * scatter: batch[dest][0..P] = (double) payload row for every row with dest >= 0 of `payload` (the rows of all ranks,
 *       gathered on the trainer rank), and batch[zero_rows[k]][:] = 0 for the all-zero first row of every chain that the
 *       RW / LG forks ship (RW:496,626). batch is fp64 [.][P + 1], the reference trainer's input layout (RV:976-985).
 *       n_payload = 0 (payload may then be NULL) writes the zero rows only. */
int sapt_train_rows_pack(const float* train_theta, const double* train_L, const int32_t* src, const int32_t* dest, int n_rows, int P,
                         float* payload, void* stream);
/* plan: the index vectors above in one launch, from the per-chain row counts train_count[R_global] (clamped to cap) and the ownership
 *       mask train_mine[R_global][cap] (one byte per entry: this rank wrote it). Valid entries of chain s are j in [skip_first, count);
 *       the batch lists the chains in order, each preceded by an all-zero row when zero_row (RW:496,626); skip_first drops entry 0
 *       (RV:683). Writes src / dest [n_pad] (-1 from n_owned on), zero_rows[R_global] (when zero_row), n_out = rows of the batch,
 *       n_owned = rows of this rank's payload that carry data. Device pointers throughout; nothing is read back. */
int sapt_train_rows_plan(const int32_t* train_count, const uint8_t* train_mine, int R_global, int cap, int n_pad, int skip_first, int zero_row,
                         int32_t* src, int32_t* dest, int64_t* zero_rows, int64_t* n_out, int64_t* n_owned, void* stream);
int sapt_train_rows_scatter(const float* payload, int64_t n_payload, int P, const int64_t* zero_rows, int n_zero, double* batch,
                            void* stream);

/* ---- chain state: scalars of the reference's ptReplica.run locals. Indexed by GLOBAL chain id
 * ("slot") except slot_of_row, which maps the rank-local theta rows to the chain each hosts.
 * Every rank keeps full-size arrays and only touches the slots its rows host; the caller
 * re-synchronises them across ranks before an exchange (a masked all-reduce of scalars). */
typedef struct {
  int32_t* slot_of_row;   /* [R local rows] global chain id hosted by that row */
  double* temperature;    /* [n_slots] T of the chain (adapttemp == temperature, RV:513-514) */
  double* lik_cur;        /* `likelihood_copy` — what MH reads (RV:586-587) */
  double* lik_post;       /* `likelihood` — what the exchange message carries (RW:623 / LG:805) */
  double* prior_cur;      /* `prior_current` (RV:588) */
  double* mov;            /* [n_slots][3] last three surg_likeh_list[:,2] entries, newest first (RV:541) */
  double* likl_wo_temp;   /* `likl_without_temp` of the last true train evaluation */
  double* last_rmse_train; double* last_rmse_test; double* last_acc_train; double* last_acc_test; /* history row i */
  int32_t* naccept;       /* [n_slots] */
} sapt_chain_state;

/* per-step history, indexed [step][slot] with slot = GLOBAL chain id; any pointer may be NULL */
typedef struct {
  int n_slots;            /* R_global */
  double* lhood_list; double* rmse_train; double* rmse_test; double* acc_train; double* acc_test;
  double* surg;           /* [step][slot][3] */
  double* likeh;          /* likeh_list[:,0] */
  double* accept_list;    /* naccept before the step (RV:578) */
  uint8_t* accepted; uint8_t* is_sur;
  float* pos_w;           /* [step][slot][P] when SAPT_F_RECORD_POS */
} sapt_history;

/* ---- Metropolis-Hastings accept + bookkeeping: replaces RV:525-642 after the evaluations ---
 * For every row: builds likelihood_proposal (true: L_train/T; surrogate: 0.5*denorm(s)/T +
 * 0.5*mean(mov), RV:540-542), mh = min(1, exp((L' - lik_cur) + (prior' - prior_cur))) with
 * overflow -> 1 (RV:571-577), u = u_inject[slot] or Philox U; on accept copies proposal -> theta
 * and commits the scalars; writes history row step+1; appends (proposal, L_untempered) of true
 * steps to the surrogate training set (RW:540-545). */
typedef struct {
  float* theta; const float* proposal; int64_t ld; int R, P;
  sapt_chain_state st;
  sapt_history hist;
  int step;                 /* i of the reference loop; history row written = step + 1 */
  uint32_t flags;           /* SAPT_F_* */
  const int32_t* is_sur;    /* [R] from sapt_propose (NULL = all true) */
  const double* L_train; const int32_t* nc_train; const double* sq_train; /* [R] from sapt_loglik */
  const double* L_test; const int32_t* nc_test; const double* sq_test;
  int64_t n_train, n_test;
  const double* prior_prop; /* [R] */
  const float* sur_s;       /* [R] surrogate outputs */
  double min_y, max_y;
  uint32_t seed; const double* u_inject; /* [n_slots] or NULL */
  /* surrogate training-set capture (true steps): rows appended per slot; NULL = off */
  float* train_theta;       /* [n_slots][train_cap][P] */
  double* train_L;          /* [n_slots][train_cap] */
  int32_t* train_count;     /* [n_slots] */
  int train_cap;
  int capture;              /* 1: capture here (RW/RV). 0: already captured by sapt_lg_first_eval */
  /* LG: the first evaluation's tempered likelihood feeds surg[:,0/2] and the window */
  const double* surg_first; /* [R] or NULL */
  const int32_t* step_dev;  /* see sapt_propose_args.step_dev */
  /* local-surrogate fork (SAPT_F_LOCAL_SURROGATE): outputs of sapt_local_surrogate_predict, [R] each */
  const double* ls_pred; const double* ls_avg;
} sapt_accept_args;
int sapt_mh_accept(const sapt_accept_args* a, void* stream);

/* ---- fused step for SMALL problems (random-walk forks): ONE launch runs n_steps complete MCMC steps of every chain
 * — proposal (Philox) + prior, likelihood on train and test, optional surrogate MLP, MH accept and bookkeeping — i.e.
 * n_steps iterations of the loop body RV:505-642 for R chains, with theta resident in shared memory. No exchange
 * happens inside; call sapt_swap_sweep between calls. `acc` is the same struct sapt_mh_accept takes (its per-row
 * input pointers, n_train/n_test and capture are ignored: the kernel computes them); acc.step (+ *acc.step_dev) is
 * the first step index. */
typedef struct {
  float* theta; int64_t ld; int R, P;
  double step_w, sigma_squared; uint32_t seed;
  int use_surrogate, surrogate_ready; double surrogate_prob; int gate;
  const float* W1; const float* b1; int h1; const float* W2; const float* b2; int h2; const float* W3; const float* b3;
  sapt_accept_args acc;
  int n_steps;
} sapt_fused_args;
int sapt_fused_steps(sapt_problem_t* p, const sapt_fused_args* a, void* stream);

/* ---- LG only: after the FIRST true evaluation (LG:622-645): capture (proposal, L) into the
 * training set, remember L/T for surg[:,0/2], throw the Langevin coin lx2 < l_prob (LG:647-649)
 * and list the rows that take a Langevin proposal. */
typedef struct {
  const float* proposal; int64_t ld; int R, P;
  sapt_chain_state st; int step; uint32_t seed;
  const int32_t* true_rows; const int32_t* n_true;
  const double* L_train;    /* [R] first evaluation */
  double* surg_first;       /* [R] out: L/T */
  int use_langevin; double l_prob; const double* lx2_inject;
  int32_t* lang_rows; int32_t* n_lang;
  float* train_theta; double* train_L; int32_t* train_count; int train_cap;
  const int32_t* step_dev;  /* see sapt_propose_args.step_dev */
} sapt_lg_first_args;
int sapt_lg_first_eval(const sapt_lg_first_args* a, void* stream);

/* ---- replica exchange: replaces ParallelTempering.swap_procedure RW:869-898 and the sequential
 * sweep RW:1095-1099 over ALL R_global chains. lik_post / temperature are the slot-indexed,
 * rank-synchronised arrays; row_of_slot[s] says which GLOBAL theta row hosts chain s.
 * The sweep runs k = 0..R-2 with a = min(1, 0.5*exp(min(709, L[k+1]-L[k]))), L = lik_post
 * (SAPT_F_POST_TEMPERED, LG:805) or lik_post*T (RW:623), u = u_inject[k] or Philox SWAP(sweep, k).
 * Whole messages swap, so the pair k+1 competes with whatever pair k left behind. Every rank
 * runs the identical sweep (counter-based uniforms), so no second collective is needed.
 * Effects: row_of_slot is re-bound (parameter vectors never move); with SAPT_F_TAKE_POST_LIK
 * lik_post[k] takes the received message's value (LG:813-815); swapped[k] = decision of pair k;
 * *num_swap += accepted pairs; src_slot[k] = chain whose message chain k received; for the
 * rank-local rows [row0, row0+R_local) slot_of_row_local is refreshed.
 * Scratch (O(R_global log R_global) bytes for long ladders) is kept per (device, stream) inside the library, behind a mutex:
 * concurrent sweeps on different streams / from different host threads do not share it. The first sweep of a given size on a
 * stream synchronises that stream and allocates, so it must not be issued under CUDA-graph capture; later ones may be. */
int sapt_swap_sweep(int R_global, double* lik_post, const double* temperature, int32_t* row_of_slot,
                    uint32_t flags, uint32_t seed, int sweep, const double* u_inject,
                    uint8_t* swapped, int32_t* num_swap, int32_t* src_slot,
                    int row0, int R_local, int32_t* slot_of_row_local, void* stream);
/* same, with the sweep index taken as sweep + *sweep_dev (graph replay) */
int sapt_swap_sweep_dev(int R_global, double* lik_post, const double* temperature, int32_t* row_of_slot,
                        uint32_t flags, uint32_t seed, int sweep, const int32_t* sweep_dev, const double* u_inject,
                        uint8_t* swapped, int32_t* num_swap, int32_t* src_slot,
                        int row0, int R_local, int32_t* slot_of_row_local, void* stream);
/* *counter += delta on the device, in stream order (advances the step / sweep counters between captured steps) */
int sapt_advance(int32_t* counter, int delta, void* stream);

/* ---- Philox4x32-10 exposed for the known-answer test */
int sapt_philox_host(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* fills u[n] (kind uniform) or z[n] (normals) on the device exactly as the kernels draw them */
int sapt_philox_fill(uint32_t seed, int kind, int step, int chain, int n, int normal, float* out, void* stream);

/* ---- n whole MCMC steps of the random-walk forks chained inside the library (no host code between the kernels):
 * the loop body RV:514-642 / RW:492-640 for all R chains, i.e. per step
 *   sapt_propose -> [sapt_surrogate_predict on the surrogate rows] -> sapt_loglik(train: all rows when the surrogate is
 *   compared or absent, RV:543, else the true rows) -> sapt_loglik(test, true rows) -> sapt_mh_accept.
 * `prop` and `acc` are the argument blocks of sapt_propose / sapt_mh_accept filled by the caller as for a single
 * step; their `step` fields give the FIRST step, the driver advances them. No exchange happens inside: call
 * sapt_swap_sweep (and the surrogate rendezvous) between calls. The Langevin fork's extra stages
 * (sapt_lg_first_eval, sapt_langevin, second proposal) are driven by the host mirror (engine.py). */
typedef struct {
  sapt_propose_args prop;
  sapt_accept_args acc;
  int use_surrogate;          /* surrogate coin thrown at all (prop.decide) */
  int compare_surrogate;      /* SAPT_F_COMPARE_SURROGATE: train likelihood on every row */
  const float* W1; const float* b1; int h1; const float* W2; const float* b2; int h2; const float* W3; const float* b3;
  float* sur_s;               /* [R] surrogate outputs (same buffer acc.sur_s reads) */
  double* L_train; int32_t* nc_train; double* sq_train;   /* [R] outputs of the train evaluation (acc reads them) */
  double* L_test; int32_t* nc_test; double* sq_test;      /* [R] */
  int32_t* counters;          /* [3] n_true, n_sur, spare: zeroed before every proposal (prop.n_true / prop.n_sur point here) */
  int path;                   /* SAPT_PATH_* */
  int n_steps;
} sapt_step_args;
int sapt_step(sapt_problem_t* p, const sapt_step_args* a, void* stream);

/* ---- N4: per-chain "local" surrogate of the fork localsurrogates/surrogate_pt_classifier_rw_localsurrogate.py (LS).
 * Every chain owns a 4 -> 10 -> 1 sigmoid network (LS:57-101) that predicts its next tempered likelihood from
 * the last four entries of surg_likeh_list[:, 2] (LS:592-598); it is retrained online at every surrogate
 * interval on lag-2 windows of five consecutive entries (LS:771-832) for 200 epochs of per-row SGD
 * (backprop_train LS:152-177). Weight layout per chain: W1[4][10] | W2[10] | B1[10] | B2 (= Network.encode order). */
#define SAPT_LS_INPUTS 4
#define SAPT_LS_HIDDEN 10
#define SAPT_LS_WEIGHTS (SAPT_LS_INPUTS * SAPT_LS_HIDDEN + 2 * SAPT_LS_HIDDEN + 1)
typedef struct {
  double* weights;             /* [n_slots][SAPT_LS_WEIGHTS] */
  double* limit;               /* [n_slots] 2 * min(y), fixed by the first training (LS:791) */
  const double* surg;          /* history surg_likeh_list [samples][n_slots][3] (sapt_history.surg) */
  int n_slots;
  const int32_t* slot_of_row;  /* [R] chains hosted by this rank */
  int R;
} sapt_local_surrogate;

/* Trains the networks of the R local chains on history rows [row_lo, row_hi) of column 2: the series is trimmed to
 * a multiple of five from the front, cut into (len-5)/2 windows with lag 2 and divided by limit (LS:773-797).
 * first != 0: limit = 2*min(series) and the weights are initialised like create_network (LS:62-65) from
 * init_inject[slot][61] standard normals (order W1, B1, W2, B2) or Philox stream LS_INIT. The regression target
 * is what the reference's BackwardPass makes of it: one_hot(int(target)) over one output = 1.0 (LS:90-93).
 * rmse[slot] (optional) = sqrt(sum of squared errors against the raw targets / windows) (LS:171-173).
 * Replaces ptReplica.local_surrogatetrainer LS:432-446 + surrogate.train LS:229-262 for all chains at once. */
int sapt_local_surrogate_train(const sapt_local_surrogate* m, int row_lo, int row_hi, int first, const double* init_inject,
                               uint32_t seed, double lrate, int epochs, double* rmse, void* stream);

/* For the listed rows (rows/n_rows as in sapt_loglik): x = surg[step-4 .. step-1][slot][2] / limit,
 * pred[row] = net(x) * limit, avg[row] = mean(x) * limit (surrogate.predict LS:313-325, LS:592-617). */
int sapt_local_surrogate_predict(const sapt_local_surrogate* m, int step, const int32_t* step_dev, const int32_t* rows,
                                 const int32_t* n_rows, double* pred, double* avg, void* stream);

/* ---- measurement: what ONE replica x row evaluation executes on the likelihood kernel `path` selects for this topology
 * (padding, split products and shared reciprocals included) — derived from the same compile-time geometry the kernels use,
 * so bench.py's roofline is computed from the instruction mix the kernel runs, not from a model of an ideal kernel. */
enum {
  SAPT_WORK_PATH = 0,             /* SAPT_PATH_F32 or SAPT_PATH_TC: what SAPT_PATH_AUTO resolves to */
  SAPT_WORK_MUFU = 1,             /* MUFU (ex2 / rcp / lg2) operations per evaluation */
  SAPT_WORK_MMA_FLOPS = 2,        /* tensor-pipe FLOPs issued per evaluation (0 on the fp32 path) */
  SAPT_WORK_FFMA = 3,             /* FP32 FMAs of the two layers per evaluation (0 on the tcgen05 path) */
  SAPT_WORK_FLOPS_ALGORITHMIC = 4, /* 2 D H + 2 H O */
  SAPT_WORK_COUNT = 5
};
int sapt_loglik_work_model(const sapt_problem_t* p, int path, double* out_host, int n_out);

/* ---- measurement: SM pipe micro-benchmarks (the denominators of bench.py's "governing" roofline; SURVEY.md 8d:
 * "pipe peaks must be re-measured by microbenchmark"). Fills out[SAPT_PIPE_COUNT]: lane-operations per second of
 * MUFU.EX2, MUFU.RCP, FFMA, FMUL, fma.rn.f32x2 (instructions x 32 lanes; one f32x2 counts once), LOP3, a 1:3
 * MUFU:FFMA mix, cvt.rn.bf16x2.f32 (F2FP, with one LOP3 per conversion), and the SM clock seen by clock64 during the FFMA / MUFU loops. Takes ~30 ms; synchronises. */
enum {
  SAPT_PIPE_SMS = 0, SAPT_PIPE_SM_HZ = 1, SAPT_PIPE_EX2 = 2, SAPT_PIPE_RCP = 3, SAPT_PIPE_FFMA = 4, SAPT_PIPE_FMUL = 5,
  SAPT_PIPE_FFMA2 = 6, SAPT_PIPE_LOP3 = 7, SAPT_PIPE_MIX = 8, SAPT_PIPE_SM_HZ_MUFU = 9, SAPT_PIPE_F2FP = 10, SAPT_PIPE_COUNT = 11
};
int sapt_measure_pipes(int device, double* out_host, int n_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SAPT_H_ */
