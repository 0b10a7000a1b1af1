// Metropolis-Hastings decision + bookkeeping for ONE chain, executed by one thread. Shared by mh_accept_kernel
// (sapt_mcmc.cu) and the fused small-problem step kernel (sapt_fused.cu) so both follow _revamp.py:525-642
// through the same code.
#pragma once
#include <math.h>

#include "sapt.h"
#include "sapt_device.cuh"

namespace sapt {

struct MhInputs {
  bool sur;                 // surrogate step
  float sur_s;              // raw surrogate output
  double L_train; int nc_train; double sq_train;
  double L_test; int nc_test; double sq_test;
  double prior_prop;
  bool has_surg_first; double surg_first;   // LG: tempered likelihood of the first evaluation
  double ls_pred, ls_avg;                   // local-surrogate fork: prediction * limit, mean(window) * limit
};

// returns accept; *cnt_out = index of the training-set row to fill with the proposal (or -1)
__device__ __forceinline__ bool mh_decide(const sapt_accept_args& a, int slot, int step_i, const MhInputs& in, int* cnt_out) {
  const int S = a.hist.n_slots;
  const double INF = __longlong_as_double(0x7ff0000000000000LL);
  const double NANV = __longlong_as_double(0x7ff8000000000000LL);
  const int64_t ho = (int64_t)(step_i + 1) * S + slot;  // history row i+1
    const double T = a.st.temperature[slot];
    const bool sur = in.sur;
    double Lprop, s0, s1, s2;
    double rmse_tr = a.st.last_rmse_train[slot], rmse_te = a.st.last_rmse_test[slot];
    double acc_tr = a.st.last_acc_train[slot], acc_te = a.st.last_acc_test[slot];
    double rm_tr_new = 0, rm_te_new = 0, ac_tr_new = 0, ac_te_new = 0;
    int cnt = -1;
    if (sur && (a.flags & SAPT_F_LOCAL_SURROGATE)) {
      // LS:598: the network's output times limit IS the (tempered) proposal likelihood: no 1/T, no window blend
      Lprop = in.ls_pred;
      double ltrue = 0.0;
      if (a.flags & SAPT_F_COMPARE_SURROGATE) {                              // LS:607-608
        ltrue = in.L_train / T;
        a.st.likl_wo_temp[slot] = in.L_train;
      }
      s0 = ltrue; s1 = Lprop; s2 = in.ls_avg;                                // LS:615-617
    } else if (sur) {
      // RV:319-320 / LG:325: float32 network output times float64 scalars
      const double s = (double)in.sur_s;
      double pred = (a.flags & SAPT_F_DENORM_HALF_MIN) ? (s * a.min_y / 2.0) : (s * (a.max_y - a.min_y) + a.min_y);
      const double surl = pred * (1.0 / T);                                  // RV:540
      const double* m = a.st.mov + (int64_t)slot * 3;
      const double mov = (m[0] + m[1] + m[2]) / 3.0;                         // RV:541
      Lprop = (surl * 0.5) + (mov * 0.5);                                    // RV:542
      double ltrue = 0.0;
      if (a.flags & SAPT_F_COMPARE_SURROGATE) {                              // RV:543-544
        ltrue = in.L_train / T;
        a.st.likl_wo_temp[slot] = in.L_train;
      }
      s0 = ltrue; s1 = Lprop; s2 = mov;                                      // RV:549-551
    } else {
      const double L = in.L_train;
      Lprop = L / T;
      s0 = Lprop; s1 = NANV; s2 = Lprop;                                     // RV:555,564-565
      if (in.has_surg_first) { s2 = in.surg_first; }                          // LG:644-645 then LG:678-679
      a.st.likl_wo_temp[slot] = L;
      rm_tr_new = sqrt(in.sq_train / (double)a.n_train);                 // RV:380-382
      ac_tr_new = 100.0 * ((double)in.nc_train / (double)a.n_train);     // RV:384-389
      if (a.n_test > 0) {
        rm_te_new = sqrt(in.sq_test / (double)a.n_test);
        ac_te_new = 100.0 * ((double)in.nc_test / (double)a.n_test);
      }
      if (a.capture && a.train_count) {                                      // RW:540-545
        cnt = a.train_count[slot];
        if (cnt < a.train_cap) { a.train_L[(int64_t)slot * a.train_cap + cnt] = L; a.train_count[slot] = cnt + 1; }
        else cnt = -1;
      }
    }
    *cnt_out = cnt;
    // moving-average window: newest first
    double* m = a.st.mov + (int64_t)slot * 3;
    m[2] = m[1]; m[1] = m[0]; m[0] = s2;
    const double prior_prop = in.prior_prop;
    const double diff = (Lprop - a.st.lik_cur[slot]) + (prior_prop - a.st.prior_cur[slot]);
    // min(1, exp(diff)); OverflowError -> 1 (RV:571-577); Python's min(1, nan) is 1 as well
    double mh;
    if (diff != diff) mh = 1.0;
    else if (diff >= 0.0) mh = 1.0;
    else mh = exp(diff);
    const double u = a.u_inject ? a.u_inject[slot] : (double)philox_uniform(a.seed, SAPT_STREAM_U, step_i, (uint32_t)slot, 0);
    const int nacc = a.st.naccept[slot];
    const bool acc = u < mh;
    double h_lhood, h_rtr, h_rte, h_atr, h_ate;
    if (acc) {
      a.st.naccept[slot] = nacc + 1;
      a.st.lik_cur[slot] = Lprop; a.st.lik_post[slot] = Lprop; a.st.prior_cur[slot] = prior_prop;
      if (!sur) { h_lhood = Lprop * T; h_rtr = rm_tr_new; h_rte = rm_te_new; h_atr = ac_tr_new; h_ate = ac_te_new; }
      else { h_lhood = h_rtr = h_rte = h_atr = h_ate = INF; }
    } else {
      if (!sur) { h_lhood = Lprop * T; h_rtr = rmse_tr; h_rte = rmse_te; h_atr = acc_tr; h_ate = acc_te; }
      else { h_lhood = h_rtr = h_rte = h_atr = h_ate = INF; }
    }
    a.st.last_rmse_train[slot] = h_rtr; a.st.last_rmse_test[slot] = h_rte;
    a.st.last_acc_train[slot] = h_atr; a.st.last_acc_test[slot] = h_ate;
    if (a.hist.lhood_list) a.hist.lhood_list[ho] = h_lhood;
    if (a.hist.rmse_train) a.hist.rmse_train[ho] = h_rtr;
    if (a.hist.rmse_test) a.hist.rmse_test[ho] = h_rte;
    if (a.hist.acc_train) a.hist.acc_train[ho] = h_atr;
    if (a.hist.acc_test) a.hist.acc_test[ho] = h_ate;
    if (a.hist.surg) { double* sg = a.hist.surg + ho * 3; sg[0] = s0; sg[1] = s1; sg[2] = s2; }
    if (a.hist.likeh) a.hist.likeh[ho] = a.st.likl_wo_temp[slot];
    if (a.hist.accept_list) a.hist.accept_list[ho] = (double)nacc;
    if (a.hist.accepted) a.hist.accepted[ho] = acc ? 1 : 0;
    if (a.hist.is_sur) a.hist.is_sur[ho] = sur ? 1 : 0;
    return acc;
}

}  // namespace sapt
